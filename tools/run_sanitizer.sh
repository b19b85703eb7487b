#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over tools/sanitizer_workload.py, plus the debug-artefact test
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_debug_dumps.py -m gpu -x -q > $O/san_pytest.txt 2>&1; tail -5 $O/san_pytest.txt
timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 python tools/sanitizer_workload.py > $O/san_mem.txt 2>&1; grep -E "ok|ERROR SUMMARY|Invalid|out of bounds|Error" $O/san_mem.txt | head -12
timeout 1500 compute-sanitizer --tool racecheck --print-limit 8 python tools/sanitizer_workload.py > $O/san_race.txt 2>&1; grep -E "ok|RACECHECK SUMMARY|hazard|Error" $O/san_race.txt | head -12
