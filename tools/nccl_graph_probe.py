"""Can a step with NCCL all-reduces be captured in a CUDA graph here?  torchrun --nproc-per-node 2 tools/nccl_graph_probe.py"""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(512, 200, device=dev) * (dist.get_rank() + 1)
y = torch.zeros_like(x)
dist.all_reduce(x.clone()); torch.cuda.synchronize()
for mode in ("thread_local", "global"):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                t = x.clone(); dist.all_reduce(t)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode=mode):
            t = x * 2.0
            dist.all_reduce(t)
            y.copy_(t + 1.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            g.replay()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 200 * 1e6
        w = dist.get_world_size()
        expect = 2.0 * sum(range(1, w + 1)) + 1.0
        print(f"rank {dist.get_rank()} mode {mode}: ok={bool((y == expect).all())} {dt:.1f} us per replay", flush=True)
    except Exception as e:
        print(f"rank {dist.get_rank()} mode {mode}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
dist.barrier(); dist.destroy_process_group()
