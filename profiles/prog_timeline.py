"""Timeline of the persistent multi-op kernel (ns_program, r01 design): where do the consumers spend an op?
Run on a B200: NS_PROG_TIMELINE=1 python profiles/prog_timeline.py [n_layers]  -> gpurun_out/prog_timeline.npz + a summary."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NS_PROG_TIMELINE", "1")
import neural_speed_b200 as ns  # noqa: E402

N_EMBD, N_FF, N_VOCAB = 4096, 11008, 32000
n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
BB = int(os.environ.get("BARRIER", "1"))
L = ns.lib()
L.bestla_init()
dev = L.bestla_create_device(False)
queue = L.bestla_get_device_queue(dev)
torch.manual_seed(0)


def make_weight(n, k):
    w = torch.randn(n, k, device="cuda") * 0.02
    rows = torch.empty(n * (k // 32) * 18, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    assert L.ns_device_quantize_q4_0(C.c_void_p(w.data_ptr()), C.c_void_p(rows.data_ptr()), n, k, queue) == 0
    h = ns.Weight.from_q4_0_device(rows.data_ptr(), n, k, k // 32 * 18, queue)
    L.bestla_device_sync(queue)
    return h


names = [("wq", N_EMBD, N_EMBD), ("wk", N_EMBD, N_EMBD), ("wv", N_EMBD, N_EMBD), ("wo", N_EMBD, N_EMBD), ("w1", N_FF, N_EMBD),
         ("w3", N_FF, N_EMBD), ("w2", N_EMBD, N_FF)]
layers = [{nm: make_weight(n, k) for nm, n, k in names} for _ in range(n_layers)]
lm_head = make_weight(N_VOCAB, N_EMBD)
x = torch.randn(1, N_EMBD, device="cuda")
attn = torch.randn(1, N_EMBD, device="cuda")
qkv = torch.zeros(3, 1, N_EMBD, device="cuda")
o = torch.zeros(1, N_EMBD, device="cuda")
tmp = torch.zeros(1, N_FF, device="cuda")
ffn = torch.zeros(1, N_EMBD, device="cuda")
logits = torch.zeros(1, N_VOCAB, device="cuda")
prog = ns.Program(1)
TAGS = int(os.environ.get("TAGS", "0"))
CHAIN = int(os.environ.get("CHAIN", "1"))  # 1: every op reads the previous op's output (the real dependency chain of a decode step)
if CHAIN:
    tq = torch.zeros(3 * N_EMBD, dtype=torch.int64, device="cuda")    # tagged copies of the outputs ({value, tag} words)
    to = torch.zeros(N_EMBD, dtype=torch.int64, device="cuda")
    tt = torch.zeros(N_FF, dtype=torch.int64, device="cuda")
    tf = torch.zeros(N_EMBD, dtype=torch.int64, device="cuda")
    tg = lambda tns: tns.data_ptr() if TAGS else None
    first = True
    for lay in layers:
        prog.add([lay["wq"], lay["wk"], lay["wv"]], ns.Program.CONCAT, (tf if TAGS and not first else ffn).data_ptr() if not first else x.data_ptr(), N_EMBD,
                 qkv.data_ptr(), 3 * N_EMBD, barrier_before=0 if first else BB, in_tagged=bool(TAGS and not first), dst_tag_ptr=tg(tq))
        first = False
        prog.add([lay["wo"]], ns.Program.PLAIN, (tq if TAGS else qkv).data_ptr(), 3 * N_EMBD, o.data_ptr(), N_EMBD, barrier_before=BB, in_tagged=bool(TAGS),
                 dst_tag_ptr=tg(to))
        prog.add([lay["w1"], lay["w3"]], ns.Program.GATE_UP_SILU, (to if TAGS else o).data_ptr(), N_EMBD, tmp.data_ptr(), N_FF, barrier_before=BB,
                 in_tagged=bool(TAGS), dst_tag_ptr=tg(tt))
        prog.add([lay["w2"]], ns.Program.PLAIN, (tt if TAGS else tmp).data_ptr(), N_FF, ffn.data_ptr(), N_EMBD, barrier_before=BB, in_tagged=bool(TAGS),
                 dst_tag_ptr=tg(tf))
    prog.add([lm_head], ns.Program.PLAIN, (tf if TAGS else ffn).data_ptr(), N_EMBD, logits.data_ptr(), N_VOCAB, barrier_before=BB, in_tagged=bool(TAGS))
else:
    for lay in layers:
        prog.add([lay["wq"], lay["wk"], lay["wv"]], ns.Program.CONCAT, x.data_ptr(), N_EMBD, qkv.data_ptr(), 3 * N_EMBD, barrier_before=BB)
        prog.add([lay["wo"]], ns.Program.PLAIN, attn.data_ptr(), N_EMBD, o.data_ptr(), N_EMBD, barrier_before=BB)
        prog.add([lay["w1"], lay["w3"]], ns.Program.GATE_UP_SILU, x.data_ptr(), N_EMBD, tmp.data_ptr(), N_FF, barrier_before=BB)
        prog.add([lay["w2"]], ns.Program.PLAIN, tmp.data_ptr(), N_FF, ffn.data_ptr(), N_EMBD, barrier_before=BB)
    prog.add([lm_head], ns.Program.PLAIN, x.data_ptr(), N_EMBD, logits.data_ptr(), N_VOCAB, barrier_before=BB)
prog.finalize(queue)
for _ in range(5):
    prog.run(queue)
L.bestla_device_sync(queue)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.ExternalStream(queue)
e0.record(st)
for _ in range(20):
    prog.run(queue)
e1.record(st)
e1.synchronize()
print("program: %.1f us/token" % (e0.elapsed_time(e1) * 1e3 / 20))
nops, grid = C.c_int(0), C.c_int(0)
buf = np.zeros(8 * 1024 * 1024, np.uint64)
L.ns_program_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
rc = L.ns_program_timeline(prog.h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nops), C.byref(grid))
assert rc == 0, ns.last_error()
KTL = 12
tl = buf[: nops.value * grid.value * KTL].reshape(nops.value, grid.value, KTL).astype(np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "prog_timeline.npz"), tl=tl)
MHZ = 1965.0
kinds = ["qkv", "o", "gate_up", "down"]
for ki, kn in enumerate(kinds + ["lm_head"]):
    idx = [nops.value - 1] if kn == "lm_head" else list(range(4 + ki, nops.value - 1, 4))
    t = tl[idx]
    wait = (t[:, :, 1] - t[:, :, 0]) / MHZ
    quant = (t[:, :, 2] - t[:, :, 1]) / MHZ
    comp = (t[:, :, 3] - t[:, :, 2]) / MHZ
    pub = (t[:, :, 4] - t[:, :, 3]) / MHZ
    prod_lead = (t[:, :, 0] - t[:, :, 5]) / MHZ  # >0: the producer had started this op's stream before the consumers got to it
    prod_span = (t[:, :, 6] - t[:, :, 5]) / MHZ
    gt = t[:, :, 7]
    spread = (gt.max(axis=1) - gt.min(axis=1)) / 1e3
    total = np.diff(tl[:, :, 7].astype(np.float64), axis=0)[[i - 1 for i in idx if i > 0]] / 1e3 if idx[0] > 0 else None
    setup = (t[:, :, 8] - t[:, :, 2]) / MHZ
    regld = (t[:, :, 9] - t[:, :, 8]) / MHZ
    print(f"{kn:8s} setup {setup.mean():5.2f} regload {regld.mean():5.2f} | wait {wait.mean():6.2f} (max-cta {wait.max(axis=1).mean():6.2f})  quant {quant.mean():5.2f}  compute {comp.mean():6.2f} "
          f"(min {comp.min(axis=1).mean():5.2f} max {comp.max(axis=1).mean():5.2f})  publish {pub.mean():5.2f}  "
          f"producer lead {prod_lead.mean():6.2f} span {prod_span.mean():6.2f}  op-start spread {spread.mean():5.2f} us")
per_op = np.diff(tl[:, 0, 7].astype(np.float64)) / 1e3
print("op-to-op (globaltimer, CTA 0) us: mean %.2f; by kind:" % per_op[4:].mean(), [round(float(per_op[4 + k::4].mean()), 2) for k in range(4)])
print("globaltimer granularity sample:", np.unique(np.diff(np.sort(tl[:, :, 7].ravel())))[:6])
# ---- per-unit trace of CTA 0 (first 8192 units of the launch)
ut = np.zeros(8192 * 8, np.uint64)
L.ns_program_unit_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
if L.ns_program_unit_trace(prog.h, ut.ctypes.data_as(C.c_void_p), ut.size) == 0:
    ut = ut.reshape(8192, 8).astype(np.int64)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "prog_units.npz"), ut=ut)
    n = int((ut[:, 2] > 0).sum())
    u = ut[300:min(n, 2000)]
    print("units traced", n)
    print("producer per unit (cycles): alloc->space %.0f, space->issued %.0f, issue-to-issue %.0f" % (
        (u[:, 1] - u[:, 0]).mean(), (u[:, 2] - u[:, 1]).mean(), np.diff(u[:, 2]).mean()))
    print("consumer per unit (cycles): wait %.0f, compute %.0f; issue->ready %.0f" % (
        (u[:, 4] - u[:, 3]).mean(), (u[:, 5] - u[:, 4]).mean(), (u[:, 4] - u[:, 2]).mean()))
    base = ut[400, 0]
    for k in range(400, 440):
        print(k, (ut[k, :6] - base).tolist())
