"""How fast can one B200 retire chains of tiny dependent kernels when K independent chains (CUDA graphs on K
streams) are in flight?  (Diagnostic for the batch-mode analysis in profiles/.)"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import time
import torch

N = 500
x = [torch.zeros(256, device="cuda") for _ in range(32)]
streams = [torch.cuda.Stream() for _ in range(32)]
graphs = []
for k in range(32):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(streams[k]):
        x[k].add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=streams[k]):
            for _ in range(N):
                x[k].add_(1.0)
    graphs.append(g)
torch.cuda.synchronize()
for K in (1, 2, 4, 8, 16, 32):
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for k in range(K):
            with torch.cuda.stream(streams[k]):
                graphs[k].replay()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print("K=%2d chains x %d tiny kernels: %.2f ms total, %.2f us per kernel per chain, %.2f us per kernel device-wide"
          % (K, N, dt * 1e3, dt * 1e6 / N, dt * 1e6 / (N * K)))
