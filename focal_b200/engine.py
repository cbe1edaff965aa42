"""Host side of the FOCAL loss hot path: shape checks, workspace, stage sequencing, row-sharded multi-GPU.

Everything numerical happens in ``libfocal_b200.so`` (hand-written sm_100a kernels behind a C ABI,
``include/focal_b200.h``); PyTorch is used for device memory, streams and ``torch.distributed`` only.
There is no CPU path and no PyTorch fallback: inputs that are not CUDA tensors raise.

Multi-GPU: R ranks each hold B/R rows (whole sequences) of every feature tensor.  The
result equals the single-GPU loss on the rank-major concatenation; every rank gets the global scalar and
d loss_global / d (its own rows).  Because logits are symmetric, a rank that knows every row sum can form
P_kj + P_jk for its own rows -- no gradient reduce-scatter is needed.  Two implementations:

* peer path (default on one NVSwitch box, ``focal_b200_loss_sharded``): the workspaces are mapped into every
  process (CUDA IPC -- or, at 8 ranks, torch symmetric memory, which adds an NVSwitch multicast mapping); the prologue
  of each rank handles its own rows and stores their bf16 operands straight into all workspaces over NVLink (one store
  per peer, or one ``multimem.st`` through the multicast mapping), row sums and loss partials travel the same way,
  three device-side barriers order the phases.  No collective library call on the data path, so the whole step is one
  CUDA graph.
* collective path (any process group; also what the CPU ``gloo`` tests drive): all-gather of the raw features
  (operands rebuilt bit-identically on every rank), all-gather of the InfoNCE row sums, all-reduce of the five
  loss partials.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _cabi


@dataclass(frozen=True)
class FocalHyper:
    """What FOCALLoss reads from args (reference loss.py:11-23,149,163,211-215)."""
    modalities: Tuple[str, ...]
    seq_len: int
    temperature: float
    margin: float
    w_shared: float
    w_private: float
    w_orth: float
    w_rank: float
    no_private: bool = False
    terms: int = _cabi.FOCAL_TERM_ALL
    precision: str = "auto"        # "bf16" | "fp32" (split-bf16 tiles, 16 significant bits) | "auto" (resolve_precision)


# "auto": batches whose step is launch-latency-sized anyway run the fp32 mode (split-bf16 tiles, three tensor-core passes
# per product); above that the Gram passes are tensor-pipe-bound and run bf16 tiles (north_star's bf16 mode).
# FOCAL_B200_PRECISION overrides "auto".
AUTO_FP32_MAX_ROWS = 2048


def resolve_precision(hp: "FocalHyper", B: int, D: int) -> int:
    import os
    want = hp.precision
    if want == "auto":
        want = os.environ.get("FOCAL_B200_PRECISION", "auto").lower()
    if want in ("fp32", "tf32"):
        return _cabi.FOCAL_PREC_FP32
    if want == "bf16":
        return _cabi.FOCAL_PREC_BF16
    if want != "auto":
        raise ValueError(f"unknown precision {want!r} (bf16 | fp32 | auto)")
    return _cabi.FOCAL_PREC_FP32 if (B <= AUTO_FP32_MAX_ROWS and D <= 256) else _cabi.FOCAL_PREC_BF16


# calls served by this process, by kind (tests use it to prove that third-party code really went through the CUDA path)
CALLS = {"grad": 0, "nograd": 0}


def shard_sequences(b: int, world: int, rank: int) -> Tuple[int, int]:
    """Sequences owned by ``rank`` when b sequences are split rank-major into equal blocks."""
    if b % world:
        raise ValueError(f"{b} sequences do not split evenly over {world} ranks (need B % (S*R) == 0)")
    per = b // world
    return rank * per, (rank + 1) * per


class CudaBackend:
    """Runs the stages of include/focal_b200.h on the current CUDA stream."""

    name = "cuda"

    def __init__(self, precision: Optional[int] = None):
        self.lib = _cabi.load()          # raises ImportError when the extension is missing -- no fallback
        self.precision = precision       # None: FocalHyper.precision decides ("auto" -> by batch size)
        self._ws: Dict[tuple, Tuple[torch.Tensor, torch.Tensor, _cabi.FocalWsInfo]] = {}
        self._plans: Dict[tuple, tuple] = {}
        self._peers: Dict[tuple, Optional[tuple]] = {}

    # -- helpers ------------------------------------------------------------------------------------
    def _cfg(self, hp: FocalHyper, B: int, D: int, need_grad: bool, seq: Tuple[int, int]) -> _cabi.FocalCfg:
        return _cabi.FocalCfg(
            B=B, S=hp.seq_len, M=len(hp.modalities), D=D, temperature=hp.temperature, margin=hp.margin,
            w_shared=hp.w_shared, w_private=hp.w_private, w_orth=hp.w_orth, w_rank=hp.w_rank,
            no_private=int(hp.no_private), need_grad=int(need_grad), terms=hp.terms,
            precision=resolve_precision(hp, B, D) if self.precision is None else self.precision,
            seq_begin=seq[0], seq_end=seq[1], num_sms=0)

    def workspace(self, cfg: _cabi.FocalCfg, device: torch.device):
        key = (cfg.B, cfg.S, cfg.M, cfg.D, cfg.no_private, cfg.terms, cfg.precision, device.index)
        hit = self._ws.get(key)
        info = _cabi.FocalWsInfo()
        rc = self.lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info))
        if rc == _cabi.FOCAL_ESHAPE:
            raise ValueError(f"focal_b200: {_cabi.strerror(rc)} (got B={cfg.B}, S={cfg.S}, M={cfg.M}, D={cfg.D}, "
                             f"T={cfg.temperature})")
        _cabi.check(rc, "focal_b200_workspace_info")
        if hit is None or hit[0].numel() < info.total_bytes + 1024:
            raw = torch.empty(info.total_bytes + 1024, dtype=torch.uint8, device=device)
            off = (-raw.data_ptr()) % 1024
            ws = raw[off: off + info.total_bytes]
            self._ws[key] = (raw, ws, info)
            hit = self._ws[key]
        return hit[1], info

    @staticmethod
    def _view(ws: torch.Tensor, off: int, nbytes: int, dtype: torch.dtype, shape) -> torch.Tensor:
        return ws[off: off + nbytes].view(dtype).view(*shape)

    def plan(self, hp: FocalHyper, B: int, D: int, need_grad: bool, seq: Tuple[int, int], dev: torch.device,
             blocked: Optional[Tuple[int, int, int]] = None, indirect: bool = False):
        """(cfg, ws, info, ws pointer, ws size) of one call configuration, cached."""
        key = (hp, B, D, need_grad, seq, dev.index, blocked, indirect)
        hit = self._plans.get(key)
        if hit is None:
            cfg = self._cfg(hp, B, D, need_grad, seq)
            cfg.indirect_ptrs = int(indirect)
            if blocked is not None:
                cfg.in_block_rows, cfg.in_block_stride = blocked[1], blocked[2]
            ws, info = self.workspace(cfg, dev)
            hit = (cfg, ws, info, C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()))
            self._plans[key] = hit
        return hit

    # -- the whole path -----------------------------------------------------------------------------
    supports_blocked = True

    def run(self, hp: FocalHyper, feats: Sequence[torch.Tensor], seq: Tuple[int, int], need_grad: bool,
            exchange_rowsum=None, blocked: Optional[Tuple[int, int, int]] = None):
        """feats: 2M fp32 CUDA tensors (view-major).  Plain mode: each is the full [B, D] tensor.  Blocked mode
        (``blocked = (B, rows_per_block, block_stride_in_floats)``): each is the first row block of a row-blocked
        tensor (the layout an all-gather of per-rank [2M, B/R, D] buffers produces).
        Returns (loss5 [5] fp32 device tensor with the partial sums of the owned rows, grads: list of 2M tensors
        holding d loss / d (owned rows) -- [rows_owned, D] -- or None)."""
        x0 = feats[0]
        D = x0.shape[1]
        B = blocked[0] if blocked is not None else x0.shape[0]
        dev = x0.device
        cfg, ws, info, wsp, wsn = self.plan(hp, B, D, need_grad, seq, dev, blocked)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        fptr = _cabi.ptr_array([t.data_ptr() for t in feats])
        loss5 = torch.empty(5, dtype=torch.float32, device=dev)
        rows0, rows1 = seq[0] * hp.seq_len, seq[1] * hp.seq_len
        grads = gptr = None
        if need_grad:
            # only the owned rows are written: hand the kernels a base pointer such that global row i lands at local
            # row i - rows0 of a [rows_owned, D] tensor
            grads = [torch.empty((rows1 - rows0, D), dtype=torch.float32, device=dev) for _ in feats]
            gptr = _cabi.ptr_array([g.data_ptr() - rows0 * D * 4 for g in grads])
        lib, ref = self.lib, C.byref(cfg)
        if exchange_rowsum is None:
            _cabi.check(lib.focal_b200_loss(ref, fptr, wsp, wsn, C.c_void_p(loss5.data_ptr()), gptr, stream),
                        "focal_b200_loss")
        else:
            _cabi.check(lib.focal_b200_prologue(ref, fptr, wsp, wsn, stream), "focal_b200_prologue")
            _cabi.check(lib.focal_b200_nce_rowsum(ref, wsp, wsn, stream), "focal_b200_nce_rowsum")
            _cabi.check(lib.focal_b200_nce_lse(ref, wsp, wsn, 0, stream), "focal_b200_nce_lse")
            if need_grad and (hp.terms & _cabi.FOCAL_TERM_NCE):
                rs = self._view(ws, info.rowsum_off, info.rowsum_bytes, torch.float32,
                                (info.n_problems, hp.seq_len, 2, info.bpad))
                exchange_rowsum(rs)            # fills in the other ranks' rows (all-gather)
                _cabi.check(lib.focal_b200_nce_lse(ref, wsp, wsn, 1, stream), "focal_b200_nce_lse(all)")
            _cabi.check(lib.focal_b200_nce_grad(ref, wsp, wsn, stream), "focal_b200_nce_grad")
            _cabi.check(lib.focal_b200_temporal(ref, wsp, wsn, stream), "focal_b200_temporal")
            _cabi.check(lib.focal_b200_finalize(ref, fptr, wsp, wsn, C.c_void_p(loss5.data_ptr()), gptr, stream),
                        "focal_b200_finalize")
        return loss5, grads

    # -- indirect mode: the launch sequence as a replayable unit (CUDA graphs) --------------------------
    def launch_indirect(self, hp: FocalHyper, B: int, D: int, need_grad: bool, seq: Tuple[int, int],
                        dev: torch.device, peer=None) -> None:
        """Enqueue every stage with cfg.indirect_ptrs = 1: the kernels take the caller's pointers from the workspace
        table (set_ptrs).  This is what FocalEngine captures into a CUDA graph."""
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if peer is None:
            cfg, ws, info, wsp, wsn = self.plan(hp, B, D, need_grad, seq, dev, None, True)
            _cabi.check(self.lib.focal_b200_loss(C.byref(cfg), None, wsp, wsn, None, None, stream),
                        "focal_b200_loss(indirect)")
        else:
            cfg = self._cfg(hp, B, D, need_grad, seq)
            cfg.local_rows, cfg.indirect_ptrs = 1, 1
            _cabi.check(self.lib.focal_b200_loss_sharded(C.byref(cfg), None, C.byref(peer[0]), C.c_size_t(peer[1]),
                                                         None, None, stream), "focal_b200_loss_sharded(indirect)")

    def set_ptrs(self, hp: FocalHyper, B: int, D: int, need_grad: bool, seq: Tuple[int, int], dev: torch.device,
                 feats: Sequence[torch.Tensor], loss5: torch.Tensor, grads: Optional[Sequence[torch.Tensor]],
                 peer=None) -> None:
        """Point the captured launch sequence at this step's inputs / outputs (one tiny launch)."""
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        fptr = _cabi.ptr_array([t.data_ptr() for t in feats])
        if peer is None:
            cfg, ws, info, wsp, wsn = self.plan(hp, B, D, need_grad, seq, dev, None, True)
            # full tensors: row i of the gradient lands at row i - rows0 of the [rows_owned, D] output
            rows0 = seq[0] * hp.seq_len
            gptr = _cabi.ptr_array([g.data_ptr() - rows0 * D * 4 for g in grads]) if grads is not None else None
        else:
            cfg = self._cfg(hp, B, D, need_grad, seq)
            cfg.local_rows, cfg.indirect_ptrs = 1, 1
            wsp, wsn = C.c_void_p(peer[0].ws[peer[0].rank]), C.c_size_t(peer[1])
            gptr = _cabi.ptr_array([g.data_ptr() for g in grads]) if grads is not None else None
        _cabi.check(self.lib.focal_b200_set_ptrs(C.byref(cfg), wsp, wsn, fptr, C.c_void_p(loss5.data_ptr()), gptr,
                                                 stream), "focal_b200_set_ptrs")

    # -- row-sharded path over NVLink peer memory ------------------------------------------------------
    @staticmethod
    def peer_eligible(hp: FocalHyper, D: int, world: int) -> bool:
        """Shapes focal_b200_loss_sharded handles (the vectorised row kernels with fused intra-sequence means)."""
        d = D // 2
        return (world <= _cabi.FOCAL_MAX_PEERS and not hp.no_private and D % 2 == 0 and d % 32 == 0 and (32 <= d <= 128 or d == 256)
                and hp.seq_len in (1, 2, 4))

    def peer_setup(self, cfg: _cabi.FocalCfg, group, dev: torch.device):
        """Collective over ``group``: allocate this rank's workspace, exchange IPC handles, map the peers'.
        Returns (FocalPeers, total_bytes, opened, own) or None when any rank could not (then every rank gets None)."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        key = (cfg.B, cfg.S, cfg.M, cfg.D, cfg.no_private, cfg.terms, cfg.precision, world, dev.index)
        if key in self._peers:
            return self._peers[key]
        lib = self.lib
        info = _cabi.FocalWsInfo()
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        ok = lib.focal_b200_workspace_info(C.byref(cfg), C.byref(info)) == 0
        # NVSwitch multicast stores: measured slower than per-peer stores at 2 GPUs (prologue 60 vs 37 us: every byte,
        # the rank's own copy included, crosses the switch at ~200 GB/s of multimem.st egress), so "auto" only takes
        # them where they cut the egress 8-fold
        mc_env = os.environ.get("FOCAL_B200_MULTICAST", "auto")
        if ok and world > 1 and (mc_env == "1" or (mc_env == "auto" and world >= 8)):
            ent = self._symm_setup(int(info.total_bytes), group, dev, world, rank)
            if ent is not None:
                self._peers[key] = ent
                return ent
        ok = ok and lib.focal_b200_peer_alloc(info.total_bytes, C.byref(own), handle) == 0
        got: List[Optional[tuple]] = [None] * world
        dist.all_gather_object(got, (bool(ok), handle.raw, int(info.total_bytes)), group=group)
        ok = all(g[0] for g in got) and len({g[2] for g in got}) == 1
        peers = _cabi.FocalPeers(rank=rank, world=world)
        opened = []
        if ok:
            for r in range(world):
                if r == rank:
                    peers.ws[r] = own.value
                    continue
                p = C.c_void_p()
                if lib.focal_b200_peer_open(got[r][1], C.byref(p)) != 0:
                    ok = False
                    break
                opened.append(p)
                peers.ws[r] = p.value
        flags: List[Optional[bool]] = [None] * world
        dist.all_gather_object(flags, bool(ok), group=group)
        if not all(flags):
            for p in opened:
                lib.focal_b200_peer_close(p)
            if own.value:
                lib.focal_b200_peer_free(own)
            self._peers[key] = None
            return None
        dist.barrier(group=group)            # nobody starts storing into a workspace that is not mapped / zeroed yet
        self._peers[key] = (peers, int(info.total_bytes), opened, own)
        return self._peers[key]

    def _symm_setup(self, total_bytes: int, group, dev: torch.device, world: int, rank: int):
        """Workspaces from torch's symmetric memory: besides the peer mappings it hands out an NVSwitch MULTICAST address
        of the same buffers, so the kernels store the operands / row sums of the owned rows once (multimem.st) instead of
        once per peer.  Collective over ``group``; None (on every rank) when any rank has no multicast support -- the
        caller then falls back to cudaMalloc + CUDA IPC."""
        import torch.distributed as dist
        t = h = None
        try:                                                # ask first: rendezvous is collective, nobody may enter it alone
            import torch.distributed._symmetric_memory as symm
            can = bool(symm._SymmetricMemory.has_multicast_support(symm.DeviceType.CUDA, dev.index))
        except Exception:
            can = False
        cans: List[Optional[bool]] = [None] * world
        dist.all_gather_object(cans, can, group=group)
        if not all(cans):
            return None
        try:
            t = symm.empty(total_bytes, dtype=torch.uint8, device=dev)
            h = symm.rendezvous(t, group.group_name)
            mine_ok = bool(h.multicast_ptr) and len(h.buffer_ptrs) == world and t.data_ptr() % 1024 == 0
        except Exception:                                   # no symmetric memory in this build / on this fabric
            mine_ok = False
        flags: List[Optional[bool]] = [None] * world
        dist.all_gather_object(flags, bool(mine_ok), group=group)
        if not all(flags):
            return None
        t.zero_()                                           # barrier epochs start at 0; padding rows are never written
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)                           # nobody stores into a workspace that is not zeroed yet
        peers = _cabi.FocalPeers(rank=rank, world=world)
        for r in range(world):
            peers.ws[r] = int(h.buffer_ptrs[r])
        peers.mc = int(h.multicast_ptr)
        return (peers, total_bytes, [], C.c_void_p(), (t, h))

    def close(self) -> None:
        """Release the peer-memory resources (IPC mappings, the cudaMalloc'ed workspaces) and the cached workspaces.
        Every rank must have finished its last step (callers barrier first): peers may still be storing into a
        workspace otherwise."""
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        for ent in self._peers.values():
            if ent is None:
                continue
            opened, own = ent[2], ent[3]
            for p in opened:
                self.lib.focal_b200_peer_close(p)
            if own.value:
                self.lib.focal_b200_peer_free(own)
        self._peers.clear()
        self._plans.clear()
        self._ws.clear()

    def run_sharded(self, hp: FocalHyper, local: Sequence[torch.Tensor], B: int, seq: Tuple[int, int], need_grad: bool,
                    peer) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
        """local: 2M fp32 CUDA tensors holding the owned rows.  Returns (GLOBAL loss5, grads of the owned rows)."""
        x0 = local[0]
        D, dev = x0.shape[1], x0.device
        cfg = self._cfg(hp, B, D, need_grad, seq)
        cfg.local_rows = 1
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        fptr = _cabi.ptr_array([t.data_ptr() for t in local])
        loss5 = torch.empty(5, dtype=torch.float32, device=dev)
        grads = gptr = None
        if need_grad:
            grads = [torch.empty_like(t) for t in local]
            gptr = _cabi.ptr_array([g.data_ptr() for g in grads])
        _cabi.check(self.lib.focal_b200_loss_sharded(C.byref(cfg), fptr, C.byref(peer[0]), C.c_size_t(peer[1]),
                                                     C.c_void_p(loss5.data_ptr()), gptr, stream),
                    "focal_b200_loss_sharded")
        return loss5, grads


class FocalEngine:
    """Validates inputs, shards rows over the process group, drives a backend.

    CUDA graphs: the launch sequence of one (shape, need_grad) is captured ONCE with ``indirect_ptrs`` -- the kernels take
    the caller's feature / gradient / loss pointers from a table in the workspace -- and every later step is
    ``set_ptrs`` + ``graph.replay()``.  Inputs may live at a new address every step (activations of a training loop
    do) and the outputs are freshly allocated tensors every step, so nothing a caller holds (autograd's saved gradients,
    ``last_parts``) is ever overwritten by a later replay, and no caller storage is pinned by the cache.
    """

    def __init__(self, hp: FocalHyper, process_group=None, backend=None, use_cuda_graph: bool = True):
        self.hp = hp
        self.group = process_group
        self.backend = backend if backend is not None else CudaBackend()
        import os
        is_cuda = getattr(self.backend, "name", "") == "cuda"
        self.use_cuda_graph = (use_cuda_graph and is_cuda and hasattr(self.backend, "launch_indirect")
                               and os.environ.get("FOCAL_B200_CUDA_GRAPH", "1") != "0")
        # row-sharded jobs: exchange over NVLink peer memory instead of collectives (one box, <= 8 ranks, CUDA backend)
        self.use_peer = (is_cuda and hasattr(self.backend, "run_sharded")
                         and os.environ.get("FOCAL_B200_PEER", "1") != "0")
        self._steps: Dict[tuple, object] = {}     # (rows, D, need_grad, device, world) -> "warm" | CUDAGraph
        self.graph_replays = 0
        self.graph_captures = 0

    # ---------------------------------------------------------------------------------------------
    def _world(self) -> Tuple[int, int]:
        if self.group is None:
            return 1, 0
        import torch.distributed as dist
        return dist.get_world_size(self.group), dist.get_rank(self.group)

    def _check(self, f1: Dict[str, torch.Tensor], f2: Dict[str, torch.Tensor]) -> List[torch.Tensor]:
        hp = self.hp
        feats: List[torch.Tensor] = []
        for f in (f1, f2):
            for m in hp.modalities:
                if m not in f:
                    raise KeyError(f"modality {m!r} missing from the feature dict")
                feats.append(f[m])
        x0 = feats[0]
        if x0.dim() != 2:
            raise ValueError(f"features must be [B, D], got {tuple(x0.shape)}")
        for t in feats:
            if t.shape != x0.shape or t.device != x0.device:
                raise ValueError("all feature tensors must share shape and device")
            if t.dtype != torch.float32:
                raise TypeError(f"features must be float32 (the reference computes in fp32), got {t.dtype}")
        if x0.shape[0] % hp.seq_len:
            # same condition under which the reference's reshape(-1, seq_len, D) raises (loss.py:154)
            raise ValueError(f"batch of {x0.shape[0]} rows is not a multiple of seq_len={hp.seq_len}")
        if getattr(self.backend, "name", "") == "cuda" and not x0.is_cuda:
            raise RuntimeError("focal_b200 runs on CUDA tensors only (no CPU fallback); move the features to the GPU")
        out = []
        for t in feats:
            t = t.detach().contiguous()
            if t.data_ptr() % 16:             # a view into the middle of a larger buffer: the kernels use 16-byte vectors
                t = t.clone()
            out.append(t)
        return out

    # ---------------------------------------------------------------------------------------------
    def loss_and_grads(self, f1: Dict[str, torch.Tensor], f2: Dict[str, torch.Tensor], need_grad: bool):
        """Returns (loss5, grads): loss5 = [total, shared, private, orth, temporal] of the GLOBAL batch,
        grads = 2M tensors shaped like the (local) inputs, or None.  Both are fresh tensors owned by the caller."""
        local = self._check(f1, f2)
        dev = local[0].device
        CALLS["grad" if need_grad else "nograd"] += 1
        if not local[0].is_cuda:
            return self._run(local, need_grad)
        with torch.cuda.device(dev):       # kernels, attributes and SM queries act on the CURRENT device
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                return self._step(local, need_grad)
            return self._run(local, need_grad)

    def _peer_of(self, local: List[torch.Tensor], need_grad: bool):
        """(B, seq, peer) of the row-sharded peer path, or None when this call does not take it."""
        world, rank = self._world()
        Bl, D = local[0].shape
        if world == 1 or not (self.use_peer and self.backend.peer_eligible(self.hp, D, world)):
            return None
        B = world * Bl
        seq = shard_sequences(B // self.hp.seq_len, world, rank)
        peer = self.backend.peer_setup(self.backend._cfg(self.hp, B, D, need_grad, seq), self.group, local[0].device)
        return None if peer is None else (B, seq, peer)

    def _step(self, local: List[torch.Tensor], need_grad: bool):
        hp, be = self.hp, self.backend
        world, _ = self._world()
        Bl, D = local[0].shape
        dev = local[0].device
        key = (Bl, D, need_grad, dev.index, world)
        st = self._steps.get(key)
        if st is None:
            # first sighting: run eagerly (loads the kernels, sets their attributes, builds workspaces and, when
            # row-sharded, runs the peer-memory handshake -- none of which may happen under stream capture)
            self._steps[key] = "warm"
            return self._run(local, need_grad)
        if st == "none":
            return self._run(local, need_grad)
        if world == 1:
            B, seq, peer = Bl, (0, Bl // hp.seq_len), None
        else:
            sh = self._peer_of(local, need_grad)
            if sh is None:                     # collective path: its all-gathers read caller memory directly
                self._steps[key] = "none"
                return self._run(local, need_grad)
            B, seq, peer = sh
        if st == "warm":
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                be.launch_indirect(hp, B, D, need_grad, seq, dev, peer)
            self._steps[key] = st = graph
            self.graph_captures += 1
        loss5 = torch.empty(5, dtype=torch.float32, device=dev)
        grads = [torch.empty_like(t) for t in local] if need_grad else None
        be.set_ptrs(hp, B, D, need_grad, seq, dev, local, loss5, grads, peer)
        st.replay()
        self.graph_replays += 1
        return loss5, grads

    def close(self) -> None:
        self._steps.clear()
        if hasattr(self.backend, "close"):
            self.backend.close()

    def _run(self, local: List[torch.Tensor], need_grad: bool):
        hp = self.hp
        world, rank = self._world()
        if world == 1:
            b = local[0].shape[0] // hp.seq_len
            loss5, grads = self.backend.run(hp, local, (0, b), need_grad, None)
            return loss5, grads
        import torch.distributed as dist
        Bl, D = local[0].shape
        nT = len(local)
        sh = self._peer_of(local, need_grad)
        if sh is not None:
            return self.backend.run_sharded(hp, local, sh[0], sh[1], need_grad, sh[2])
        # (1) all-gather the raw features: per-rank [2M, Bl, D] -> [R, 2M, Bl, D]
        mine = torch.stack(local, dim=0)
        gathered = torch.empty((world * nT, Bl, D), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.group)       # concatenated along dim 0, rank-major
        b = world * Bl // hp.seq_len
        seq = shard_sequences(b, world, rank)
        per = seq[1] - seq[0]

        def exchange_rowsum(rs: torch.Tensor):
            # rs: [P, S, 2, bpad]; every rank computed the k-range [seq0, seq1): all-gather the slices, one permuted copy back
            part = rs[..., seq[0]:seq[1]].contiguous()
            allp = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
            dist.all_gather_into_tensor(allp, part, group=self.group)
            P_, S_ = part.shape[0], part.shape[1]
            rs[..., :b].view(P_, S_, 2, world, per).copy_(allp.view(world, P_, S_, 2, per).permute(1, 2, 3, 0, 4))

        if getattr(self.backend, "supports_blocked", False):
            # kernels read the gathered buffer in place: row i of tensor t sits in block i // Bl
            heads = [gathered[t] for t in range(nT)]
            loss5, grads = self.backend.run(hp, heads, seq, need_grad, exchange_rowsum,
                                            blocked=(world * Bl, Bl, nT * Bl * D))
        else:
            full = [gathered.view(world, nT, Bl, D)[:, t].reshape(world * Bl, D) for t in range(nT)]
            loss5, gfull = self.backend.run(hp, full, seq, need_grad, exchange_rowsum)
            r0 = seq[0] * hp.seq_len
            grads = [g[r0: r0 + Bl] if g.shape[0] != Bl else g for g in gfull] if need_grad else None
        # (3) loss partials of the owned rows -> global loss on every rank
        dist.all_reduce(loss5, op=dist.ReduceOp.SUM, group=self.group)
        return loss5, grads
