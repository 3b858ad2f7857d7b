"""per-node timing of the integer tensor-core batch path: 8 distinct weights of one shape cycled (64+ MB > ... not L2 resident for the
big shapes), captured in one CUDA graph; prints us per node and the HBM fraction"""
import sys, os, ctypes as C, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import neural_speed_b200 as ns
L = ns.lib(); L.bestla_init()
devh = L.bestla_create_device(False)
queue = L.bestla_get_device_queue(devh)
stream = torch.cuda.ExternalStream(queue)
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6586.0) if os.path.exists("MEASURED_PEAKS.json") else 6586.0
cp = lambda t: C.c_void_p(t.data_ptr())
def capture(fn):
    assert L.ns_graph_begin(queue) == 0, ns.last_error()
    fn()
    g = L.ns_graph_end(queue); assert g, ns.last_error()
    return C.c_void_p(g)
def timed(g, reps=20):
    for _ in range(3): L.ns_graph_launch(g, queue)
    L.bestla_device_sync(queue); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): L.ns_graph_launch(g, queue)
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps
modes = {"q4_0": dict(group=32, stype=ns.S_F16, comp=ns.COMP_Q8_0, asym=False),
         "g128_asym_bf16": dict(group=128, stype=ns.S_BF16, comp=ns.COMP_INT8, asym=True),
         "g128_sym_f32": dict(group=128, stype=ns.S_F32, comp=ns.COMP_INT8, asym=False)}
only = sys.argv[1:] or list(modes)
NW = 16
for mode in only:
    kw = modes[mode]
    shapes = [(4096, 4096), (12288, 4096), (4096, 11008), (22016, 4096), (32000, 4096)]
    if os.environ.get("IMMA_SHAPES"):
        shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["IMMA_SHAPES"].split(",")]
    for (n, k) in shapes:
        ws = [ns.Weight.random(n, k, seed=3 + i, queue=queue, **kw) for i in range(NW)]
        L.bestla_device_sync(queue)
        bytes_node = ws[0].algorithmic_bytes
        for M in [int(v) for v in os.environ.get("IMMA_MS", "1,4,8,16,32").split(",")]:
            x = torch.randn(M, k, device="cuda"); y = torch.zeros(M, n, device="cuda")
            wsb = torch.zeros(L.ns_device_workspace_bytes(M, k), dtype=torch.uint8, device="cuda")
            def calls():
                for w in ws:
                    assert L.ns_mul_mat(w.h, cp(x), k, cp(y), n, M, None, None, 0, cp(wsb), queue) == 0, ns.last_error()
            g = capture(calls)
            ms = timed(g)
            us = ms * 1000 / NW
            print(f"{mode:16s} n={n:6d} k={k:6d} M={M:3d}: {us:8.2f} us/node  {bytes_node / us / 1e3 / peak * 100:5.1f}% of HBM peak", flush=True)
        del ws
