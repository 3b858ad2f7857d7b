mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_program.py -x -q 2>&1 | tail -5
for v in "A=1" "TAGS=1" "BARRIER=0" "TAGS=1 NS_PROG_GROUP_KB=13"; do echo "== $v"; env $v timeout 200 python profiles/prog_timeline.py 2>&1 | tail -56 | head -8; done
