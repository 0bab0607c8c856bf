#!/bin/bash
scripts/quick_variants.sh "--no-single-frame" "--opt gather_chunks=4" "--opt gather_chunks=8" "--opt gather_chunks=16" "--opt gather_chunks=8 --opt gather_cluster_size=8" "--opt gather_chunks=32"
