import sys, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import _lib
L = _lib.load(); sink = torch.zeros(4, device='cuda')
for name, fn, per in (("FFMA", L.cw_ffma_peak, 64), ("FFMA2", L.cw_ffma2_peak, 128)):
    for threads in (128, 256, 512):
        blocks = 148 * (2048 // threads); iters = 20000
        fn(blocks, threads, 10, sink.data_ptr(), None); torch.cuda.synchronize()
        best = 0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(blocks, threads, iters, sink.data_ptr(), None); e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * blocks * threads * iters * per / e0.elapsed_time(e1) / 1e9)
        print(f"{name} threads={threads} (64 warps/SM): {best:.1f} TFLOP/s")
    for threads, bps in ((256, 2),):
        blocks = 148 * bps; iters = 40000
        fn(blocks, threads, 10, sink.data_ptr(), None); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(blocks, threads, iters, sink.data_ptr(), None); e1.record(); torch.cuda.synchronize()
        print(f"{name} 16 warps/SM (the scoring kernel's occupancy): {2.0 * blocks * threads * iters * per / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s")
