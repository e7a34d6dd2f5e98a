#!/bin/bash
# Round-end validation on the GPU box: GPU test suite, smoke, bench (both arms), ncu launch list of one c5 step and a full capture
# of the dB epilogue.  Outputs under gpurun_out/<tag>_*.
tag=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; tail -c 300 gpurun_out/${tag}_ref.json
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_c5.csv python tools/one_step.py c5 > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:db_epilogue -o gpurun_out/${tag}_epilogue python tools/one_step.py c5 > /dev/null 2>&1
ls -la gpurun_out/${tag}_*
