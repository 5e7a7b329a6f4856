"""BatchNorm+SiLU backward (reduce + apply + param) and forward apply on the largest YOLOX-s layer shapes at batch 64, CUDA-event timed.
   YB200_BN_RED=U:MINB:ITERS python tools/bench_bn.py     (reduce-kernel variant; one process per setting)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from yolov7_d2_b200 import capi

L = capi.lib()
dev = torch.device("cuda:0")
shapes = [(64, 320, 320, 32), (64, 160, 160, 64), (64, 160, 160, 32), (64, 80, 80, 128), (64, 80, 80, 64), (64, 40, 40, 256), (64, 40, 40, 128)]
tot_ms, tot_b = 0.0, 0.0
print("setting", os.environ.get("YB200_BN_RED", "default"))
for n, h, w, c in shapes:
    z = torch.randn(n, h, w, c, device=dev).to(torch.float16)
    da = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
    dz = torch.empty_like(da)
    scale, shift = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
    mean, invstd = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    a1, a2 = torch.zeros(c, dtype=torch.float64, device=dev), torch.zeros(c, dtype=torch.float64, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    za, daa, dza = capi.act(z), capi.act(da), capi.act(dz)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run():
        capi.check(L.yb200_bn_silu_bwd(ctypes.byref(za), ctypes.byref(daa), None, None, capi.ptr(scale), capi.ptr(shift), capi.ptr(mean), capi.ptr(invstd),
                                       capi.ptr(a1), capi.ptr(a2), ctypes.byref(dza), capi.ptr(dg), capi.ptr(db), 0, capi.stream_ptr()), "bn_silu_bwd")

    for _ in range(3):
        run()
    ts = []
    for _ in range(5):
        flush.zero_()  # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 10.0 * n * h * w * c
    tot_ms += ms; tot_b += nbytes
    print("  %dx%dx%dx%-4d reduce+apply+param %.3f ms  %.0f GB/s (10 B/element actual traffic)" % (n, h, w, c, ms, nbytes / ms / 1e6))
print("  total %.3f ms  %.0f GB/s" % (tot_ms, tot_b / tot_ms / 1e6))
