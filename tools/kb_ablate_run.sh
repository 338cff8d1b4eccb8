#!/bin/bash
# perf_probe on the production library and on each experiment variant under lib/exp/
PKG=jmd-diversity-in-bayesian-optimization_b200
export NO_BO=1 NS=${NS:-16,32,48,64,96,110}
echo "== base"; python tools/perf_probe.py 2>&1 | grep kbl_frac | sed 's/.*n.: \([0-9]*\),.*kbl_ms.: \([0-9.]*\),.*kbl_frac.: \([0-9.]*\).*/n=\1 ms=\2 frac=\3/'
for f in $PKG/lib/exp/libb200bo_*.so; do
  echo "== $f"; B200BO_LIB_PATH=$PWD/$f python tools/perf_probe.py 2>&1 | grep kbl_frac | sed 's/.*n.: \([0-9]*\),.*kbl_ms.: \([0-9.]*\),.*kbl_frac.: \([0-9.]*\).*/n=\1 ms=\2 frac=\3/'
done
