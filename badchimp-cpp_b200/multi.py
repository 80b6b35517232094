"""N ranks, one process per GPU: z-slab decomposition with halo exchange over torch.distributed.

Replaces MonLatMpi::communicateLbField (src/lbsolver/LBmonlatmpi.h:236-297): the engine packs
the outgoing populations of the two z-faces on the device (haloPackKernel), this module moves
the packed buffers with NCCL send/recv on the engine's halo stream while the interior nodes
are still being collided, and the engine unpacks them into the halo-in slots.

The z direction is periodic, so the ranks form a ring: face "down" talks to rank-1, face "up"
to rank+1.  With 2 ranks both faces talk to the same peer; messages between one pair of ranks
are matched in posting order, so sends are posted (down, up) and receives (up, down): the
peer's "down" message is what arrives at my "up" face.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist


class RingHalo:
    """send/recv buffers of the two z-faces and the exchange itself (works on CPU/gloo and CUDA/NCCL)"""

    def __init__(self, rank, world, n_send_down, n_recv_down, n_send_up, n_recv_up, device):
        self.rank, self.world = rank, world
        self.down, self.up = (rank - 1) % world, (rank + 1) % world
        mk = lambda n: torch.zeros(max(int(n), 1), dtype=torch.float64, device=device)
        self.send_down, self.recv_down = mk(n_send_down), mk(n_recv_down)
        self.send_up, self.recv_up = mk(n_send_up), mk(n_recv_up)
        self.counts = (int(n_send_down), int(n_recv_down), int(n_send_up), int(n_recv_up))

    def ops(self):
        sd, rd, su, ru = self.counts
        return [dist.P2POp(dist.isend, self.send_down[:sd], self.down),
                dist.P2POp(dist.isend, self.send_up[:su], self.up),
                dist.P2POp(dist.irecv, self.recv_up[:ru], self.up),
                dist.P2POp(dist.irecv, self.recv_down[:rd], self.down)]

    def exchange(self):
        reqs = dist.batch_isend_irecv(self.ops())
        for r in reqs:
            r.wait()


def attach_ring(lat, slab, rank, world, device):
    """registers the two halo faces of a structured-ingest lattice and wires the NCCL exchange"""
    faces = slab["faces"]
    lat.add_halo_face((rank - 1) % world, faces["down"][0].cpu().numpy(), faces["down"][1].cpu().numpy())
    lat.add_halo_face((rank + 1) % world, faces["up"][0].cpu().numpy(), faces["up"][1].cpu().numpy())
    lat.set_boundary_count(slab["n_boundary"])
    nf = lat.n_fields   # a packed message holds the face of every LbField, field after field
    ring = RingHalo(rank, world, nf * len(faces["down"][0]), nf * len(faces["down"][1]), nf * len(faces["up"][0]), nf * len(faces["up"][1]), device)
    lat.set_halo_buffers(0, ring.send_down.data_ptr(), ring.recv_down.data_ptr())
    lat.set_halo_buffers(1, ring.send_up.data_ptr(), ring.recv_up.data_ptr())

    def on_exchange(stream_ptr):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            ring.exchange()

    lat.set_exchange_callback(on_exchange)
    lat._ring = ring
    return ring


def attach_ring_peer(lat, slab, rank, world):
    """registers the two halo faces and connects them to the neighbour GPUs' memory (CUDA IPC): the engine
    then stores outgoing populations straight into the peers' halo-in slots, no NCCL and no host callback
    per step.  torch.distributed only carries the one-time handshake (IPC handles + receive lists)."""
    faces = slab["faces"]
    down, up = (rank - 1) % world, (rank + 1) % world
    lat.add_halo_face(down, faces["down"][0].cpu().numpy(), faces["down"][1].cpu().numpy())
    lat.add_halo_face(up, faces["up"][0].cpu().numpy(), faces["up"][1].cpu().numpy())
    lat.set_boundary_count(slab["n_boundary"])
    info = {"handles": lat.ipc_handles(), "field_stride": lat.nq * lat.plane_stride(),
            "recv_down": faces["down"][1].cpu().numpy(), "recv_up": faces["up"][1].cpu().numpy()}
    everyone = [None] * world
    dist.all_gather_object(everyone, info)
    # what I send down lands in the "up" face (index 1) of rank-1, what I send up in the "down" face (0) of rank+1
    lat.connect_peer(0, everyone[down]["field_stride"], 1, everyone[down]["recv_up"], handles=everyone[down]["handles"])
    lat.connect_peer(1, everyone[up]["field_stride"], 0, everyone[up]["recv_down"], handles=everyone[up]["handles"])
    dist.barrier()


def attach_ring_twophase(lat, slab, rank, world, device):
    """two-field lattice on a z-slab: population halos (attach_ring), the scalar halo of phi between the moment
    and the collide pass (communciateScalarField, main_TWOPHASE.cpp:287) and the all-reduce of the momentum sum
    (:299), all over torch.distributed (NCCL on GPUs) on the stream the engine hands to the callbacks."""
    ring = attach_ring(lat, slab, rank, world, device)
    sf = slab["scalar_faces"]
    sring = RingHalo(rank, world, len(sf["down"][0]), len(sf["down"][1]), len(sf["up"][0]), len(sf["up"][1]), device)
    lat.add_scalar_halo_face(0, sf["down"][0].cpu().numpy(), sf["down"][1].cpu().numpy(), sring.send_down.data_ptr(), sring.recv_down.data_ptr())
    lat.add_scalar_halo_face(1, sf["up"][0].cpu().numpy(), sf["up"][1].cpu().numpy(), sring.send_up.data_ptr(), sring.recv_up.data_ptr())
    cuda = device.type == "cuda"
    scratch = torch.zeros(16, dtype=torch.float64, device=device)
    rt = C.CDLL("libcudart.so") if cuda else None
    if rt is not None:
        rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def on_scalar(stream_ptr):
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            sring.exchange()

    def on_allreduce(dev_ptr, count, stream_ptr):
        # the engine's scalar lives in its own allocation: stage it through a torch tensor on the same stream
        assert count <= scratch.numel()
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            assert rt.cudaMemcpyAsync(scratch.data_ptr(), dev_ptr, 8 * count, 3, stream_ptr) == 0
            dist.all_reduce(scratch[:count])
            assert rt.cudaMemcpyAsync(dev_ptr, scratch.data_ptr(), 8 * count, 3, stream_ptr) == 0

    lat.set_scalar_exchange_callback(on_scalar)
    lat.set_allreduce_callback(on_allreduce)
    lat._sring, lat._scratch = sring, scratch
    return ring


def attach_ring_twophase_peer(lat, slab, rank, world):
    """two-field lattice on a z-slab with every exchange over peer memory (CUDA IPC + NVLink): population faces
    and phi faces are stored straight into the neighbours' arrays, the momentum sum goes through per-rank
    mailboxes and is added in rank order on every GPU.  No NCCL call and no host callback inside a step;
    torch.distributed only carries the one-time handshake."""
    faces, sf = slab["faces"], slab["scalar_faces"]
    down, up = (rank - 1) % world, (rank + 1) % world
    lat.add_halo_face(down, faces["down"][0].cpu().numpy(), faces["down"][1].cpu().numpy())
    lat.add_halo_face(up, faces["up"][0].cpu().numpy(), faces["up"][1].cpu().numpy())
    lat.add_scalar_halo_face(0, sf["down"][0].cpu().numpy(), sf["down"][1].cpu().numpy())
    lat.add_scalar_halo_face(1, sf["up"][0].cpu().numpy(), sf["up"][1].cpu().numpy())
    lat.set_boundary_count(slab["n_boundary"])
    h2 = lat.ipc_handles_twophase()
    info = {"handles": lat.ipc_handles(), "phi": h2[:64], "mail": h2[64:], "field_stride": lat.nq * lat.plane_stride(),
            "recv_down": faces["down"][1].cpu().numpy(), "recv_up": faces["up"][1].cpu().numpy(),
            "phi_recv_down": sf["down"][1].cpu().numpy(), "phi_recv_up": sf["up"][1].cpu().numpy()}
    everyone = [None] * world
    dist.all_gather_object(everyone, info)
    # what I send down lands in the "up" face (index 1) of rank-1, what I send up in the "down" face (0) of rank+1
    lat.connect_peer(0, everyone[down]["field_stride"], 1, everyone[down]["recv_up"], handles=everyone[down]["handles"])
    lat.connect_peer(1, everyone[up]["field_stride"], 0, everyone[up]["recv_down"], handles=everyone[up]["handles"])
    lat.connect_peer_scalar(0, everyone[down]["phi_recv_up"], handle=everyone[down]["phi"])
    lat.connect_peer_scalar(1, everyone[up]["phi_recv_down"], handle=everyone[up]["phi"])
    lat.connect_world(rank, world, handles=[e["mail"] for e in everyone])
    dist.barrier()


def peer_memory_available(local, world, device):
    """True when every GPU of this node can address every other one (NVLink / NVSwitch, or PCIe peer-to-peer):
    the peer-memory transport needs it; otherwise callers fall back to the NCCL transport.  The answer is
    all-reduced so that all ranks take the same branch."""
    ok = 1.0
    try:
        n = torch.cuda.device_count()
        if n < world:
            ok = 0.0
        else:
            for other in range(world):
                if other != local and not torch.cuda.can_device_access_peer(local, other):
                    ok = 0.0
    except Exception:
        ok = 0.0
    t = torch.tensor([ok], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() == 1.0)


def balanced_cuts(my_layer_counts, world):
    """z cut planes such that every rank owns (nearly) the same number of fluid nodes: my_layer_counts holds the
    fluid nodes of each z-layer of this rank's equal-thickness slab (int64 tensor on the rank's device)"""
    all_layers = [torch.zeros_like(my_layer_counts) for _ in range(world)]
    dist.all_gather(all_layers, my_layer_counts)
    cum = torch.cumsum(torch.cat(all_layers), 0).cpu().numpy()
    targets = cum[-1] * np.arange(1, world) / world
    cuts = [0]
    for t in targets:
        k = int(np.searchsorted(cum, t))          # cum[k-1] < t <= cum[k]
        below = cum[k - 1] if k > 0 else 0
        cuts.append(k if (t - below) < (cum[k] - t) else k + 1)
    cuts.append(len(cum))
    return cuts


def run_weak_scaling(args, pkg, ingest, size, lattice, tau, force):
    """bench.py --gpus N (N > 1): every rank owns one size^3 block of a size x size x (size N) pack"""
    from . import bench_impl as B
    capi = pkg.capi
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    t_setup = time.perf_counter()
    gshape = (size, size, size * world)
    z0, z1 = rank * size, (rank + 1) * size
    ext = ingest.sphere_pack_slab(gshape, size / 8.0, 0.35, 1234, z0 - 1, z1 + 1)
    if getattr(args, "balance", True):
        # cut planes chosen so that every rank owns the same number of fluid nodes (the step time is the
        # maximum over ranks; equal-thickness slabs of a random pack differ by several per cent)
        layers = torch.from_numpy(ext[:, :, 1:-1].reshape(-1, size).sum(axis=0).astype(np.int64)).to(device)
        cuts = balanced_cuts(layers, world)
        z0, z1 = cuts[rank], cuts[rank + 1]
        ext = ingest.sphere_pack_slab(gshape, size / 8.0, 0.35, 1234, z0 - 1, z1 + 1)
    slab = ingest.build_slab_tables(torch.from_numpy(ext).to(device).bool(), lattice, boundary_first=True)
    index_form = capi.INDEX_COMPACT if args.index == "compact" else capi.INDEX_TABLE
    lat = capi.lattice_from_device_table(lattice, slab["n"], slab["n_pad"], slab["n_halo"], slab["table"].data_ptr(),
                                         slab["labels"].data_ptr(), 1, index_form, local)
    halo_mode = getattr(args, "halo", "peer")
    if halo_mode == "peer" and not peer_memory_available(local, world, device):
        halo_mode = "nccl"   # no peer addressing between the GPUs of this node
    if halo_mode == "peer":
        attach_ring_peer(lat, slab, rank, world)
        halo_bytes = 8.0 * (len(slab["faces"]["down"][0]) + len(slab["faces"]["up"][0]))
    else:
        attach_ring(lat, slab, rank, world, device)
        halo_bytes = 8.0 * (sum(lat._ring.counts[0::2]))
    n = slab["n"]
    slab["table"] = slab["labels"] = None
    torch.cuda.empty_cache()
    lat.init_uniform(1.0)
    setup_s = time.perf_counter() - t_setup

    def total(x, op=dist.ReduceOp.SUM):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        return float(t.item())

    n_total = total(n)
    lat.step_single(args.warmup, tau=tau, force=force)
    lat.synchronize()
    samples, stop = [], threading.Event()
    th = threading.Thread(target=B._clock_sampler, args=(stop, samples), daemon=True)
    if rank == 0:
        th.start()
    dist.barrier()
    torch.cuda.synchronize()
    l1 = capi.lib().chimp_launch_count()
    ms = lat.step_timed(args.steps, tau=tau, force=force)
    lat.synchronize()
    l2 = capi.lib().chimp_launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    ms_max = total(ms, dist.ReduceOp.MAX)
    per_rank = torch.zeros(world, 3, dtype=torch.float64, device=device)
    per_rank[rank] = torch.tensor([float(n), float(ms) / args.steps, float(z1 - z0)], dtype=torch.float64, device=device)
    dist.all_reduce(per_rank)
    per_rank = per_rank.cpu().numpy()
    if rank == 0:
        t_extra = time.perf_counter()
        while len(samples) < 6 and time.perf_counter() - t_extra < 2.0:
            time.sleep(0.1)
    # keep all ranks in step while rank 0 samples clocks under load
    lat.step_single(20, tau=tau, force=force)
    lat.synchronize()
    stop.set()
    rho, _ = lat.download_moments_device_order()
    mass_err = abs(total(rho.sum()) / n_total - 1.0)

    # end to end per rank: upload LbField (pinned host, reference AoS rows 0..N), K steps, download rho + vel
    e2e = None
    try:
        nq = len(pkg.geometry.BASIS[lattice])
        host_f = torch.empty((n + 1, nq), dtype=torch.float64, pin_memory=True)
        host_f[:] = torch.from_numpy(pkg.cases.lattice_weights(lattice))[None, :]
        host_rho = torch.empty((n + 1,), dtype=torch.float64, pin_memory=True)
        host_vel = torch.empty((n + 1, 3), dtype=torch.float64, pin_memory=True)
        lib = capi.lib()
        p = lat._single_params(tau, force, None)
        # one untimed cycle first (first-touch of the pinned pages by the copy engines), like the warm-up steps
        capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
        capi._check(lib.chimp_step_single(lat.h, C.byref(p), C.c_int(1)))
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        capi._check(lib.chimp_upload_lbfield(lat.h, C.c_void_p(host_f.data_ptr())))
        capi._check(lib.chimp_step_single(lat.h, C.byref(p), C.c_int(args.steps)))
        capi._check(lib.chimp_download_rho(lat.h, C.c_void_p(host_rho.data_ptr()), C.c_int(1)))
        capi._check(lib.chimp_download_vel(lat.h, C.c_void_p(host_vel.data_ptr())))
        dt = total(time.perf_counter() - t0, dist.ReduceOp.MAX)
        e2e = {"value": n_total * args.steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": total(host_f.numel() * 8) / args.steps,
               "d2h_bytes_per_step": total((host_rho.numel() + host_vel.numel()) * 8) / args.steps,
               "cycle": "per rank: upload LbField (pinned host, reference AoS) + %d steps with halo exchange + download rho, vel; wall clock, max over ranks" % args.steps}
    except Exception as exc:  # pragma: no cover
        e2e = {"value": None, "unit": "MLUPS", "error": str(exc)}

    if rank == 0:
        peak, peak_src = B._peaks()
        kernel_ms = ms_max / args.steps
        achieved = B.B_ALG[lattice] * n_total / world / (kernel_ms * 1e-3) / 1e9
        line = {"metric": "MLUPS", "value": n_total * args.steps / (ms_max * 1e-3) / 1e6, "unit": "MLUPS", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "std_case physics (D3Q19 BGK + Guo force + half-way bounce back), periodic random sphere pack %dx%dx%d (one %d^3 block per GPU, z-slabs), R=%d, seed 1234" % (size, size, size * world, size, size // 8),
                           "fluid_nodes": n_total, "index_form": args.index, "parallelism": "z-slab x%d, halos overlapped with interior nodes" % world,
                           "l2_policy": "state per GPU 2 x %.1f GB >> 126 MB L2" % (n * 152 / 1e9),
                           "halo_bytes_per_step_per_gpu": halo_bytes,
                           "halo_transport": "peer stores over NVLink from the halo-coupled part of the step (CUDA IPC)" if halo_mode == "peer" else "NCCL send/recv (torch.distributed) between pack and unpack kernels",
                           "slabs": "balanced by fluid-node count" if getattr(args, "balance", True) else "equal thickness",
                           "nodes_per_rank": [int(x) for x in per_rank[:, 0]], "ms_per_step_per_rank": [round(float(x), 4) for x in per_rank[:, 1]],
                           "slab_thickness_per_rank": [int(x) for x in per_rank[:, 2]],
                           "setup_seconds": setup_s, "mean_rho_error": mass_err},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": B._traffic_gb(lattice, args.index, n_total / world)[0], "traffic_unit": "GB per launch per GPU (ncu dram read+write)",
                             "peak_source": peak_src, "per": "GPU (mean)"},
                "e2e": e2e, "gpu_launches": int(l2 - l1), "clocks": B._summarize_clocks(samples)}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
