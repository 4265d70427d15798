#!/bin/bash
II2_BUCKET=576 bash scratch/sweep.sh "-DK1B_CAP_N=768 -DK1B_MIN_CTAS=5" "-DK1B_CAP_N=768 -DK1B_MIN_CTAS=4"
II2_BUCKET=832 bash scratch/sweep.sh "-DX_BASE"
