"""The vendor's TF32 GEMM under the same ncu metrics as the ASPP kernels (comparison point for `sm__pipe_tensor_cycles_active`
and `sm__mem_tensor_cycles_active` of kind::tf32 tiles fed from shared memory):

    ncu --set full --clock-control none --profile-from-start off -o gpurun_out/cublas_tf32 python tools/cublas_tf32_probe.py

One torch.matmul (fp32 storage, allow_tf32) of 8192 x 8192 x 8192 and one of the ASPP GEMM shape (65536 x 18432 x 256) between
cudaProfilerStart / Stop.  Also prints CUDA-event timings (outside the profiled region) when run without ncu."""
import sys
import torch

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device('cuda:0')
shapes = [(8192, 8192, 8192), (65536, 18432, 256)]
ops = []
for m, k, n in shapes:
    a = torch.randn(m, k, device=dev)
    b = torch.randn(k, n, device=dev)
    ops.append((a, b, 2.0 * m * k * n))
for a, b, fl in ops:
    for _ in range(3):
        a @ b
torch.cuda.synchronize()
if '--time' in sys.argv:
    for (a, b, fl), s in zip(ops, shapes):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print('cuBLAS tf32 %s: %.3f ms  %.1f TFLOP/s' % (s, ms, fl / ms / 1e9))
torch.cuda.cudart().cudaProfilerStart()
for a, b, fl in ops:
    a @ b
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
