#!/bin/bash
# GPU tests only; log under gpurun_out/$1
O=gpurun_out/${1:-tests}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -25 $O/pytest.log
