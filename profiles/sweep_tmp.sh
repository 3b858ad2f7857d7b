mkdir -p gpurun_out
for v in "NS_PROG_INFLIGHT=15" "NS_PROG_INFLIGHT=20" "NS_PROG_BATCH=8 NS_PROG_INFLIGHT=8" "NS_PROG_BATCH=8 NS_PROG_INFLIGHT=12" "NS_PROG_BATCH=10 NS_PROG_INFLIGHT=10" "NS_PROG_BATCH=30" "NS_PROG_SMEM_KB=150 NS_PROG_INFLIGHT=15"; do echo "== $v"; env $v timeout 200 python profiles/prog_timeline.py 2>&1 | tail -56 | head -11; done
