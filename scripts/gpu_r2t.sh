#!/bin/bash
# Round 2, GPU call T: the round-end sequence as the driver runs it - GPU suite (-x), smoke(), bench.py.
set -u
TAG=r02t
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
