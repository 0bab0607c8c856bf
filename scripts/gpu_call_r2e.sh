#!/bin/bash
scripts/quick_variants.sh "--no-single-frame" "" "--opt gather_shared_batches=6" "--opt gather_shared_batches=12" "--opt gather_shared_batches=1000" "--opt gather_shared_batches=12 --opt gather_cluster_skip_max=4" "--opt gather_shared_batches=1000 --opt gather_cluster_skip_max=0" "--opt gather_vpl_batches=8"
