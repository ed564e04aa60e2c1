"""Developer probe (torchrun, 2+ GPUs): NCCL send/recv bandwidth between ranks 0 and 1, and the
copy-engine peer-copy bandwidth between devices 0 and 1 measured inside rank 0."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
peer = rank ^ 1
for it in range(4):
    torch.cuda.synchronize(); dist.barrier(); t = time.time()
    if peer < world:
        ops = [dist.P2POp(dist.isend, a, peer), dist.P2POp(dist.irecv, b, peer)]
        for w in dist.batch_isend_irecv(ops): w.wait()
    torch.cuda.synchronize(); dt = time.time() - t
    if rank == 0: print(f"nccl sendrecv 512 MiB each way: {dt*1e3:.2f} ms -> {n/dt/1e9:.1f} GB/s per direction", flush=True)
# the exchange of one column: many pieces per peer in one group
for npieces in (8, 80, 240):
    sz = n // npieces
    for it in range(3):
        torch.cuda.synchronize(); dist.barrier(); t = time.time()
        ops = []
        for k in range(npieces):
            ops.append(dist.P2POp(dist.isend, a[k * sz:(k + 1) * sz], peer)); ops.append(dist.P2POp(dist.irecv, b[k * sz:(k + 1) * sz], peer))
        for w in dist.batch_isend_irecv(ops): w.wait()
        torch.cuda.synchronize(); dt = time.time() - t
    if rank == 0: print(f"{npieces} pieces, 512 MiB total each way: {dt*1e3:.2f} ms -> {n/dt/1e9:.1f} GB/s", flush=True)
if rank == 0 and torch.cuda.device_count() > 1:
    x = torch.empty(n, dtype=torch.uint8, device="cuda:0"); y = torch.empty(n, dtype=torch.uint8, device="cuda:1")
    print("can access peer:", torch.cuda.can_device_access_peer(0, 1))
    for it in range(3):
        torch.cuda.synchronize(0); torch.cuda.synchronize(1); t = time.time()
        y.copy_(x, non_blocking=True); torch.cuda.synchronize(0); torch.cuda.synchronize(1); dt = time.time() - t
        print(f"peer copy 512 MiB: {dt*1e3:.2f} ms -> {n/dt/1e9:.1f} GB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
