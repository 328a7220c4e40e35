#!/bin/bash
for f in opticomlib_b200/_var_*.so; do
  echo "== $f"; SSFM_B200_LIB=$PWD/$f python scripts/exp_fused.py 1024 2>&1 | grep -E "LL|fixed"
done
