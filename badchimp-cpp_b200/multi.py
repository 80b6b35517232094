"""N ranks, one process per GPU: z-slab decomposition, halo transports and their one-time set-up.

Replaces MonLatMpi::communicateLbField / communicateScalarField and the MPI_Allreduce of the flux controller
(src/lbsolver/LBmonlatmpi.h:181-297, twophase/main_TWOPHASE.cpp:299).  Two transports:

* peer memory (default, attach_ring_peer / attach_ring_twophase_peer): the ranks exchange CUDA IPC handles and receive
  lists once over torch.distributed; afterwards the engine stores halos straight into the neighbours' memory over NVLink
  and synchronises through arrival counters -- for single-field lattices from inside the step kernel (one launch per
  step) -- with no NCCL call and no host callback per step;
* NCCL (attach_ring / attach_ring_twophase): the engine packs the outgoing populations of the two z-faces
  (haloPackKernel), this module moves the packed buffers with send/recv on the engine's halo stream while the interior
  nodes are still being collided, and the engine unpacks them into the halo-in slots.

The z direction is periodic, so the ranks form a ring: face "down" talks to rank-1, face "up" to rank+1.  With 2 ranks
both faces talk to the same peer; NCCL messages between one pair of ranks are matched in posting order, so sends are
posted (down, up) and receives (up, down): the peer's "down" message is what arrives at my "up" face.
The bench harness that uses all this is bench_impl.py / workloads.py.
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


def balanced_cuts(my_layer_counts, world, thickness=None):
    """z cut planes such that every rank owns (nearly) the same number of fluid nodes: my_layer_counts holds the
    fluid nodes of each z-layer of this rank's slab (int64 tensor on the rank's device, the same length on every
    rank; thickness[k] = number of valid leading entries of rank k when the slabs are not equally thick)"""
    all_layers = [torch.zeros_like(my_layer_counts) for _ in range(world)]
    dist.all_gather(all_layers, my_layer_counts)
    if thickness is not None:
        all_layers = [a[: int(t)] for a, t in zip(all_layers, thickness)]
    cum = torch.cumsum(torch.cat(all_layers), 0).cpu().numpy()
    targets = cum[-1] * np.arange(1, world) / world
    cuts = [0]
    for t in targets:
        k = int(np.searchsorted(cum, t))          # cum[k-1] < t <= cum[k]
        below = cum[k - 1] if k > 0 else 0
        cuts.append(k if (t - below) < (cum[k] - t) else k + 1)
    cuts.append(len(cum))
    for k in range(1, len(cuts)):                 # every rank keeps at least one layer
        cuts[k] = max(cuts[k], cuts[k - 1] + 1)
    return cuts


def attach_allreduce(lat, device):
    """MPI_Allreduce(SUM) of a few doubles in place on the device (mass change per interior domain,
    std_one_phase/main.cpp:528) over torch.distributed, on the stream the engine hands to the callback"""
    scratch = torch.zeros(64, dtype=torch.float64, device=device)
    cuda = device.type == "cuda"
    rt = C.CDLL("libcudart.so") if cuda else None
    if rt is not None:
        rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

    def on_allreduce(dev_ptr, count, stream_ptr):
        assert count <= scratch.numel()
        with torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr)):
            assert rt.cudaMemcpyAsync(scratch.data_ptr(), dev_ptr, 8 * count, 3, stream_ptr) == 0
            dist.all_reduce(scratch[:count])
            assert rt.cudaMemcpyAsync(dev_ptr, scratch.data_ptr(), 8 * count, 3, stream_ptr) == 0

    lat.set_allreduce_callback(on_allreduce)
    lat._scratch_allreduce = scratch
