"""Host-side plumbing of multi-GPU runs (one process per GPU, torch.distributed for rendezvous only).

The reference is single-process; everything here is new capability for the cut-off path's spatial domain
decomposition (include/mdsea_b200.h: rank / world / proc_grid, comm_init).  The data path between GPUs is
inside the CUDA library (NCCL send/recv + allreduce, csrc/comm.cu); this module only
  * picks the process grid,
  * hands every rank the NCCL unique id,
  * writes the per-rank compact dataset rows into the shared simfile (every particle is owned by exactly one
    rank, so the ranks write disjoint columns of r_coords / v_coords; rank 0 writes the scalar datasets),
  * reassembles full (ndim, N) host arrays from the ranks' owned pieces.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np


def dist_info() -> Tuple[int, int]:
    """(rank, world) of the current process: torch.distributed if initialised, else the torchrun env, else (0, 1)."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device: int, sysfs: str = "/sys/bus/pci/devices") -> Optional[dict]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function).
    Pinned host buffers allocated afterwards are first-touched -- and therefore placed -- on that node, so the D2H row
    traffic of N ranks does not cross the socket interconnect or pile up on one node's memory controllers.  Returns
    {"pci": ..., "numa_node": ..., "cpus": n} or None when the topology cannot be read (no change then)."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(device)
        pci = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = os.path.join(sysfs, pci)
        with open(os.path.join(base, "local_cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        node = -1
        try:
            with open(os.path.join(base, "numa_node")) as f:
                node = int(f.read().strip())
        except OSError:
            pass
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return {"pci": pci, "numa_node": node, "cpus": len(allowed), "changed": False}
        os.sched_setaffinity(0, cpus)
        return {"pci": pci, "numa_node": node, "cpus": len(cpus), "changed": True}
    except Exception:
        return None


def choose_proc_grid(world: int, ncells: Optional[int] = None) -> Tuple[int, int, int]:
    """Process grid (px, py, pz) with px*py*pz == world for the spatial decomposition of the cell grid.

    x is the fastest-running cell index: particles of an x-row of cells are contiguous in memory and the list kernels
    stream whole rows.  The grid therefore splits y and z and leaves x whole (1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2,
    8 -> 1x2x4: at 8 ranks the halo surface equals that of 2x2x2, with 5 instead of 7 distinct neighbour ranks and no faces
    that cut the rows), choosing the most square py <= pz.  x is split only if y and z alone cannot hold `world`
    (`ncells`, cells per box edge, bounds every factor: a rank needs at least one cell per dimension)."""
    best = None
    for px in range(1, world + 1):
        if world % px:
            continue
        for py in range(1, world // px + 1):
            if (world // px) % py:
                continue
            pz = world // px // py
            if py > pz:
                continue
            if ncells is not None and max(px, py, pz) > ncells:
                continue
            score = (px, pz - py)
            if best is None or score < best[0]:
                best = (score, (px, py, pz))
    if best is None:
        raise ValueError(f"no process grid for world={world} with {ncells} cells per edge")
    return best[1]


def broadcast_unique_id(make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the 128-byte NCCL id, every rank returns it (object broadcast: works on gloo and nccl)."""
    import torch.distributed as dist

    rank, world = dist_info()
    box = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id broadcast failed")
    return bytes(uid)


def barrier() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def scatter_rows(sm, rows: dict, nsteps: int, first_step: int, rank: int) -> None:
    """Write one batch of per-rank compact rows into the shared datasets.

    rows: r / v (nsteps, ndim, width), ids (nsteps, width), n_owned (nsteps,), pe / ke / temp (nsteps,).
    Column j < n_owned[k] of row k belongs to particle ids[k, j] (include/mdsea_b200.h, mdsea_rows)."""
    rds = sm.dfile[sm.rcoord_dsname]
    vds = sm.dfile[sm.vcoord_dsname]
    r_raw = getattr(rds, "raw", None)
    v_raw = getattr(vds, "raw", None)
    for k in range(nsteps):
        m = int(rows["n_owned"][k])
        ids = np.asarray(rows["ids"][k, :m])
        order = np.argsort(ids, kind="stable")  # ascending columns: friendlier to the page cache / h5py selections
        ids = ids[order]
        if r_raw is not None:
            r_raw[first_step + k][:, ids] = rows["r"][k][:, :m][:, order]
            v_raw[first_step + k][:, ids] = rows["v"][k][:, :m][:, order]
        else:  # real h5py: fancy selection along the last axis
            rds[first_step + k, :, ids] = rows["r"][k][:, :m][:, order]
            vds[first_step + k, :, ids] = rows["v"][k][:, :m][:, order]
    if rank == 0:
        sm.update_ds_block(sm.potenergy_dsname, np.asarray(rows["pe"][:nsteps]), first_step)
        sm.update_ds_block(sm.kinenergy_dsname, np.asarray(rows["ke"][:nsteps]), first_step)
        sm.update_ds_block(sm.temp_dsname, np.asarray(rows["temp"][:nsteps]), first_step)


def assemble_owned(arr: np.ndarray) -> np.ndarray:
    """Full (ndim, N) array from per-rank arrays that hold NaN outside the owned entries (allreduce of the pieces)."""
    import torch
    import torch.distributed as dist

    rank, world = dist_info()
    if world == 1:
        return arr
    piece = np.where(np.isnan(arr), 0.0, arr)
    t = torch.from_numpy(np.ascontiguousarray(piece))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()
