#!/bin/bash
# Round-end style check on a GPU box: GPU tests, smoke(), the bench line and the reference arm; logs under gpurun_out/.
tag=${1:-check}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?" >> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_smoke.log; tail -1 gpurun_out/${tag}_bench.err; head -c 300 gpurun_out/${tag}_bench.json
