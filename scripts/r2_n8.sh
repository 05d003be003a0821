#!/bin/bash
TAG=${1:-r2n8}
bash scripts/r2_multi.sh $TAG 8 20
bash scripts/r2_sweep.sh $TAG "8 4"
