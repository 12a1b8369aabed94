#!/bin/bash
# r02i: whole GPU suite + full bench line with everything of this round in place
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_tests.log
tail -12 gpurun_out/r02i_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02i_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02i_smoke.log
