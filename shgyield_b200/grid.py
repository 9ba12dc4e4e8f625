"""
Rectangular sweeps (energy x theta x phi x thickness x 4 polarisations) on one or several B200s.

The reference can only express a sweep by broadcasting NumPy arrays through `shg.shgyield`
(reference shg.py:381-437; e.g. energy (E,), theta (T,1), phi (P,1,1)).  `shgyield_grid` is the
additive entry point for the large cases of BASELINE.json (configs 3-5): it takes the four axes
explicitly and returns the yield cube in the kernel's layout [T, D, 4, P, E] (pol = pp, ps, sp, ss;
energy contiguous).

Multi-GPU: every point is independent, so theta (or thickness) is cut into contiguous shards, one
per rank (one process per GPU, torch.distributed).  Each rank's slab is a contiguous block of the
cube; the only collective is the optional gather of the slabs (NCCL over NVLink).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .shg import _physics, _prepare, physical_constants

POLS = ("pp", "ps", "sp", "ss")


def shard_bounds(n: int, world: int, rank: int) -> tuple:
    """Contiguous shard [lo, hi) of range(n) for `rank` of `world`: the first n % world ranks get
    one extra element (181 thetas over 8 ranks -> 23 x 5 + 22 x 3)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def choose_shard_axis(nT: int, nD: int, world: int) -> str:
    """Shard theta when there are enough thetas for every rank, else thickness if that has."""
    if nT >= world or nD < world:
        return "theta"
    return "thick"


def _problem(energy, theta, phi, thick, eps1w, eps2w, chi2, phys, t_range=None, d_range=None, keep=None):
    """Builds the sshg_problem struct; `keep` collects the arrays whose memory it points to."""
    def host(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a

    def addr(a, dt):
        if isinstance(a, np.ndarray) or not hasattr(a, "data_ptr"):
            return host(a, dt).ctypes.data
        keep.append(a)
        return a.data_ptr()

    def size(a):
        return int(a.numel()) if hasattr(a, "numel") else int(np.asarray(a).size)

    nE, nT, nP, nD = size(energy), size(theta), size(phi), size(thick)
    t0, t1 = t_range if t_range is not None else (0, nT)
    d0, d1 = d_range if d_range is not None else (0, nD)
    return _lib.Problem(nE=nE, nT=nT, nP=nP, nD=nD,
                        energy_eV=addr(energy, np.float64), theta_deg=addr(theta, np.float64),
                        phi_deg=addr(phi, np.float64), thick_nm=addr(thick, np.float64),
                        eps1w=addr(eps1w, np.complex128), eps2w=addr(eps2w, np.complex128),
                        chi2=addr(chi2, np.complex128), phys=phys, t0=t0, t1=t1, d0=d0, d1=d1)


def yield_grid_host(energy, theta, phi, thick, eps1w, eps2w, chi2, mref=True, consts=None, sinc_normalized=True,
                    zzz_phi_quirk=True, t_range=None, d_range=None, out=None, pinned=False, sigma_out=0.0,
                    strict_broadening=False):
    """Host arrays in, host cube out: float64 [T_shard, D_shard, 4, P, E].  The library computes
    slabs on the device and copies slab k back while slab k+1 is computed.  sigma_out > 0 broadens
    every yield spectrum along the energy axis (shg.py:433-436) inside the yield kernel (K3; row segments next
    to an azimuthal node are re-evaluated row-wise by the fix-up kernel); strict_broadening=True filters every
    row itself (yield kernel -> Gaussian FIR, ~4.5x slower)."""
    ctx = _lib.context()
    keep = []
    flags = _lib.FLAG_STRICT_BROADENING if strict_broadening else 0
    p = _problem(energy, theta, phi, thick, eps1w, eps2w, chi2,
                 _physics(consts, mref, sinc_normalized, zzz_phi_quirk, flags), t_range, d_range, keep)
    shape = (p.t1 - p.t0, p.d1 - p.d0, 4, p.nP, p.nE)
    if keep[4].shape != (3, p.nE) or keep[5].shape != (3, p.nE) or keep[6].shape != (18, p.nE):
        raise ValueError("eps1w/eps2w must be [3, nE] and chi2 [18, nE]")
    if out is None:
        out = _lib.pinned_empty(shape) if pinned else np.empty(shape)
    elif out.shape != shape or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be a C-contiguous float64 array of shape %s" % (shape,))
    if out.size == 0:
        return out                                  # empty shard (more ranks than thetas)
    _lib.check(ctx.lib.sshg_yield_broadened(ctx.handle, C.byref(p), float(sigma_out), out.ctypes.data, _lib.HOST),
               ctx.lib)
    return out


def yield_grid_device(energy, theta, phi, thick, eps1w, eps2w, chi2, out=None, mref=True, consts=None,
                      sinc_normalized=True, zzz_phi_quirk=True, t_range=None, d_range=None, stream=None,
                      sigma_out=0.0, out_ptr=None, strict_broadening=False):
    """Device-resident sweep: every array is a CUDA torch tensor (float64 / complex128), `out` a
    float64 CUDA tensor [T_shard, D_shard, 4, P, E] (allocated if None).  Asynchronous: the kernels
    are queued on torch's current stream (or `stream`).  sigma_out > 0: yields broadened along the
    energy axis in the same pass (K3).  `out_ptr` (instead of `out`): a raw device address the shard is
    written to -- e.g. `PeerCube.slab_ptr()`, the cube in another GPU's memory (returns None then)."""
    import torch
    ctx = _lib.context(energy.device.index)
    s = stream if stream is not None else torch.cuda.current_stream(energy.device)
    keep = []
    for t, dt in ((energy, torch.float64), (theta, torch.float64), (phi, torch.float64), (thick, torch.float64),
                  (eps1w, torch.complex128), (eps2w, torch.complex128), (chi2, torch.complex128)):
        if not (t.is_cuda and t.dtype == dt and t.is_contiguous()):
            raise ValueError("device sweep needs contiguous CUDA tensors of float64 / complex128")
    flags = _lib.FLAG_STRICT_BROADENING if strict_broadening else 0
    p = _problem(energy, theta, phi, thick, eps1w, eps2w, chi2,
                 _physics(consts, mref, sinc_normalized, zzz_phi_quirk, flags), t_range, d_range, keep)
    shape = (p.t1 - p.t0, p.d1 - p.d0, 4, p.nP, p.nE)
    if tuple(eps1w.shape) != (3, p.nE) or tuple(eps2w.shape) != (3, p.nE) or tuple(chi2.shape) != (18, p.nE):
        raise ValueError("eps1w/eps2w must be [3, nE] and chi2 [18, nE]")
    if out_ptr is not None:
        if out is not None:
            raise ValueError("pass either out or out_ptr")
        if int(out_ptr) % 16:
            raise ValueError("out_ptr must be 16-byte aligned")
        dst = int(out_ptr)
    else:
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=energy.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float64 or not out.is_contiguous() or not out.is_cuda:
            raise ValueError("out must be a contiguous float64 CUDA tensor of shape %s" % (shape,))
        dst = out.data_ptr()
    if out is not None and out.numel() == 0:
        return out                                  # empty shard (more ranks than thetas): nothing to compute
    if p.t1 == p.t0 or p.d1 == p.d0:
        return out
    # queued on the caller's stream; the context returns to its own stream afterwards (ordered on the device), so it
    # never keeps the handle of a stream the caller may drop, and host-array calls keep running on the own stream
    with ctx.on_stream(s.cuda_stream):
        _lib.check(ctx.lib.sshg_yield_broadened(ctx.handle, C.byref(p), float(sigma_out), dst, _lib.DEVICE), ctx.lib)
    return out


def shgyield_grid(energy, eps_m1, eps_m2, eps_m3, chi2, theta, phi, thick, gamma=90, sigma_eps=0.0, sigma_chi=0.0,
                  sigma_out=0.0, mode="3-layer", mref=True, device=True, shard=None, gather=False, strict=False):
    """
    Sweep with the reference's material arguments (same meaning as shg.shgyield) over explicit 1-D
    axes energy[E], theta[T], phi[P], thick[D].  Returns {'energy','theta','phi','thick',
    'cube': [T, D, 4, P, E], 'pp','ps','sp','ss': views [T, D, P, E]} -- torch CUDA tensors when
    `device` (default), else NumPy arrays.

    shard=(rank, world[, 'theta'|'thick']) computes only that rank's slab; with gather=True and an
    initialised torch.distributed process group the slabs are gathered onto rank 0 (`gather_slabs`, NCCL);
    gather="peer" (theta shards on one node) instead lets every rank's kernel store its slab directly into
    rank 0's cube over NVLink (`PeerCube`): no slab buffer, no collective.
    sigma_out > 0 broadens along the energy axis only (a sweep has no meaning for blurring angles); it is
    done inside the yield kernel (K3: the cube is written once; the row segments next to an azimuthal node are
    re-evaluated row-wise, so the result holds the reference to 1e-10 element-wise).  strict=True broadens
    every row itself instead (K2 -> K1 by slabs, ~4.5x slower).
    """
    energy = np.atleast_1d(np.asarray(energy, dtype=np.float64)).ravel()
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64)).ravel()
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64)).ravel()
    thick = np.atleast_1d(np.asarray(thick, dtype=np.float64)).ravel()
    consts = physical_constants()
    eps1w, eps2w, chi_rot, thick_eff = _prepare(energy, eps_m1, eps_m2, eps_m3, chi2, thick, gamma, sigma_eps,
                                                sigma_chi, mode)
    thick = np.atleast_1d(np.asarray(thick_eff, dtype=np.float64)).ravel()
    t_range = d_range = None
    if shard is not None:
        rank, world = int(shard[0]), int(shard[1])
        axis = shard[2] if len(shard) > 2 else choose_shard_axis(theta.size, thick.size, world)
        if axis == "theta":
            t_range = shard_bounds(theta.size, world, rank)
        else:
            d_range = shard_bounds(thick.size, world, rank)
    if device and gather == "peer" and t_range is not None:
        # the gather disappears into the kernel: every rank stores its theta slab straight into rank 0's cube
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", _lib.context().device)
        put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)
        dargs = (put(energy), put(theta), put(phi), put(thick), put(eps1w), put(eps2w), put(chi_rot))
        try:
            pc = PeerCube((theta.size, thick.size, 4, phi.size, energy.size), dst=0)     # raises on every rank or on none
        except RuntimeError:
            pc = None                               # no CUDA IPC / peer access on this node: compute, then NCCL gather
        if pc is not None:
            yield_grid_device(*dargs, mref=mref, consts=consts, t_range=t_range, sigma_out=sigma_out,
                              out_ptr=pc.slab_ptr(t_range[0]), strict_broadening=strict)
            torch.cuda.synchronize()
            dist.barrier()
            cube = pc.tensor()
            pc.close()
        else:
            cube = yield_grid_device(*dargs, mref=mref, consts=consts, t_range=t_range, sigma_out=sigma_out,
                                     strict_broadening=strict)
            cube = gather_slabs(cube, theta.size, thick.size, shard)
    elif device:
        import torch
        dev = torch.device("cuda", _lib.context().device)
        put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)
        cube = yield_grid_device(put(energy), put(theta), put(phi), put(thick), put(eps1w), put(eps2w), put(chi_rot),
                                 mref=mref, consts=consts, t_range=t_range, d_range=d_range, sigma_out=sigma_out,
                                 strict_broadening=strict)
        if gather and shard is not None:          # True, or "peer" on a thickness shard (not contiguous in the cube)
            cube = gather_slabs(cube, theta.size, thick.size, shard)
    else:
        cube = yield_grid_host(energy, theta, phi, thick, eps1w, eps2w, chi_rot, mref=mref, consts=consts,
                               t_range=t_range, d_range=d_range, sigma_out=sigma_out, strict_broadening=strict)
    res = {"energy": energy, "theta": theta, "phi": phi, "thick": thick, "cube": cube}
    for k, pol in enumerate(POLS):
        res[pol] = cube[:, :, k] if cube is not None else None     # gather: only rank 0 holds the cube
    return res


class _RawCuda:
    """CUDA array interface over a raw device allocation (keeps its owner alive)."""

    def __init__(self, ptr, shape, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerCube:
    """The full yield cube [T, D, 4, P, E] in the HBM of rank `dst`, mapped into every rank of the node.

    north_star's one exchange step is the gather of the theta slabs onto one GPU.  Here the destination
    rank allocates the cube (through the library, so that it has a CUDA IPC handle), the handle travels
    through torch.distributed, and every other rank maps the cube into its address space.  A rank then
    passes `slab_ptr(t_lo)` as the output of the yield kernel: its 128-bit stores go over NVLink /
    NVSwitch straight into the destination's memory -- the gather needs no slab buffer, no send/recv and
    no extra pass.  Collective calls (every rank must take part): the constructor and `close`.
    """

    def __init__(self, shape, dst: int = 0):
        import torch.distributed as dist
        self.shape = tuple(int(x) for x in shape)
        self.rank, self.dst = dist.get_rank(), int(dst)
        self.ctx = _lib.context()
        self.owner = self.rank == self.dst
        self.ptr = None
        nbytes = 8 * int(np.prod(self.shape, dtype=np.int64))
        import torch
        box, err = [None], None
        if self.owner:
            try:
                p = C.c_void_p()
                _lib.check(self.ctx.lib.sshg_device_alloc(self.ctx.handle, C.byref(p), nbytes), self.ctx.lib)
                self.ptr = int(p.value)
                h = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
                _lib.check(self.ctx.lib.sshg_ipc_export(self.ctx.handle, C.c_void_p(self.ptr), h), self.ctx.lib)
                box[0] = h.raw
            except Exception as e:              # the peers are waiting in the broadcast: tell them
                err = e
        dist.broadcast_object_list(box, src=self.dst)
        if box[0] is not None and not self.owner:
            try:
                p = C.c_void_p()
                _lib.check(self.ctx.lib.sshg_ipc_open(self.ctx.handle, box[0], C.byref(p)), self.ctx.lib)
                self.ptr = int(p.value)
            except Exception as e:
                err = e
        # every rank must agree before anybody stores through the mapping
        ok = torch.tensor([0 if (err is not None or self.ptr is None) else 1], device=torch.device("cuda", self.ctx.device))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self._release()
            raise RuntimeError("PeerCube: mapping the cube of rank %d failed on some rank%s"
                               % (self.dst, (": %r" % (err,)) if err is not None else ""))

    def slab_ptr(self, t_lo: int) -> int:
        """Address of theta row `t_lo` of the cube (valid on every rank)."""
        return self.ptr + 8 * int(t_lo) * int(np.prod(self.shape[1:], dtype=np.int64))

    def tensor(self):
        """The cube as a torch tensor on the destination rank (None elsewhere); the tensor keeps the
        allocation alive."""
        if not self.owner:
            return None
        import torch
        return torch.as_tensor(_RawCuda(self.ptr, self.shape, self), device=torch.device("cuda", self.ctx.device))

    def close(self):
        """Peers unmap the cube (after their queued kernels finished); the owner's memory lives until the
        tensor / this object is dropped."""
        if not self.owner:
            self._release()

    def _release(self):
        if not self.ptr:
            return
        if self.owner:
            self.ctx.lib.sshg_device_free(self.ctx.handle, C.c_void_p(self.ptr))
        else:
            self.ctx.lib.sshg_ipc_close(self.ctx.handle, C.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


def to_host(tensor) -> np.ndarray:
    """A contiguous float64 CUDA tensor as a fresh NumPy array, copied by the library's threaded staged copy
    (sshg_download: page-locked staging buffers, several host threads take the first-touch page faults) -- about seven
    times faster into new memory than tensor.cpu().numpy() for a multi-GB cube."""
    import torch
    if not (tensor.is_cuda and tensor.dtype == torch.float64 and tensor.is_contiguous()):
        raise ValueError("to_host needs a contiguous float64 CUDA tensor")
    out = np.empty(tuple(tensor.shape))
    if out.size == 0:
        return out
    torch.cuda.current_stream(tensor.device).synchronize()
    ctx = _lib.context(tensor.device.index)
    _lib.check(ctx.lib.sshg_download(ctx.handle, out.ctypes.data, tensor.data_ptr(), out.nbytes), ctx.lib)
    return out


def broaden_energy_device(cube, sigma):
    """K1 along the energy (last, contiguous) axis of a device cube; returns a new tensor."""
    import torch
    ctx = _lib.context(cube.device.index)
    out = torch.empty_like(cube)
    n = cube.shape[-1]
    if cube.numel() == 0:
        return out
    rows = cube.numel() // n
    with ctx.on_stream(torch.cuda.current_stream(cube.device).cuda_stream):
        _lib.check(ctx.lib.sshg_gauss1d(ctx.handle, cube.data_ptr(), out.data_ptr(), rows, n, n, 1, float(sigma),
                                        _lib.DEVICE), ctx.lib)
    return out


def gather_slabs(slab, nT: int, nD: int, shard, dst: int = 0, out=None, all_ranks: bool = False):
    """The one collective of the path: gathers the per-rank slabs into the full cube [T, D, 4, P, E].

    Default: a gather onto rank `dst` with the true (uneven) shard sizes -- grouped point-to-point
    send / recv (ncclSend / ncclRecv under NCCL, over NVLink), received straight into the slices of
    the full cube, no padding and no staging copy for theta shards (they are contiguous blocks).
    Returns the cube on `dst` and None elsewhere.  all_ranks=True gives every rank the cube
    (one broadcast per shard).  Works on CUDA tensors (NCCL) and CPU tensors (gloo, CPU tests)."""
    import torch
    import torch.distributed as dist
    rank, world = int(shard[0]), int(shard[1])
    axis = shard[2] if len(shard) > 2 else choose_shard_axis(nT, nD, world)
    n_axis = nT if axis == "theta" else nD
    dim = 0 if axis == "theta" else 1
    bounds = [shard_bounds(n_axis, world, r) for r in range(world)]
    full_shape = list(slab.shape)
    full_shape[dim] = n_axis
    if slab.shape[dim] != bounds[rank][1] - bounds[rank][0]:
        raise ValueError("slab extent %d does not match shard %s" % (slab.shape[dim], bounds[rank]))
    want_full = all_ranks or rank == dst
    full = None
    if want_full:
        if out is not None:
            if list(out.shape) != full_shape or out.dtype != slab.dtype or not out.is_contiguous():
                raise ValueError("out must be a contiguous tensor of shape %s" % (full_shape,))
            full = out
        else:
            full = torch.empty(full_shape, dtype=slab.dtype, device=slab.device)
        lo, hi = bounds[rank]
        mine = full.narrow(dim, lo, hi - lo)
        if mine.data_ptr() != slab.data_ptr():
            mine.copy_(slab)
    if all_ranks:
        for r, (lo, hi) in enumerate(bounds):
            if hi == lo:
                continue
            view = full.narrow(dim, lo, hi - lo)
            buf = view if view.is_contiguous() else view.contiguous()
            dist.broadcast(buf, src=r)
            if buf.data_ptr() != view.data_ptr():
                view.copy_(buf)
        return full
    ops, staged = [], []
    if rank == dst:
        for r, (lo, hi) in enumerate(bounds):
            if r == dst or hi == lo:
                continue
            view = full.narrow(dim, lo, hi - lo)
            buf = view if view.is_contiguous() else torch.empty(view.shape, dtype=slab.dtype, device=slab.device)
            if buf is not view:
                staged.append((view, buf))
            ops.append(dist.P2POp(dist.irecv, buf, r))
    elif slab.shape[dim] > 0:
        ops.append(dist.P2POp(dist.isend, slab.contiguous(), dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for view, buf in staged:
        view.copy_(buf)
    return full
