#!/bin/bash
O=gpurun_out/${1:-mb}; mkdir -p $O
timeout 600 python tools/mbconv_check.py --batch 8 --times-batch 256 > $O/mbconv_check.log 2>&1; echo "exit $?" >> $O/mbconv_check.log
tail -40 $O/mbconv_check.log
