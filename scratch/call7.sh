#!/bin/bash
for b in 384 448; do echo "II2_BUCKET=$b"; II2_BUCKET=$b bash scratch/sweep.sh "-DK1B_THREADS_N=128 -DK1B_CAP_N=512 -DK1B_MIN_CTAS=8" "-DK1B_CAP_N=512 -DK1B_MIN_CTAS=6" "-DK1B_CAP_N=512 -DK1B_MIN_CTAS=5"; done
