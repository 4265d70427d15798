#!/bin/bash
for sw in 592 1184 2368; do for sb in 128 256; do echo "II2_SMALL_WANT=$sw II2_SMALL_BUCKET=$sb"; II2_SMALL_WANT=$sw II2_SMALL_BUCKET=$sb python scratch/c3prof.py 2>&1 | tail -6 | grep -v "^frac 0.1\|^frac 0.001" ; done; done
