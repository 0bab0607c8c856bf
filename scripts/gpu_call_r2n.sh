#!/bin/bash
scripts/quick_variants.sh "--no-single-frame" "--opt shaft_leaf_max=1" "--opt shaft_leaf_max=3"
