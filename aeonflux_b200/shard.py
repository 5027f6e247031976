"""Multi-GPU sharding and mixed-shape streaming for the batch engine (BASELINE config 5).

Every presentation / issuance is an independent unit of work (the reference's CompactProof forbids cross-item aggregation,
/root/reference/src/nizk/presentation.rs:31), so the path shards with no data-path collective:

  * one process per GPU, each holding a replicated issuer context (`Issuer`);
  * rank k of G takes the contiguous item range [k*N/G, (k+1)*N/G) of every batch;
  * only the accept/reject bitmap (1 bit per item) is gathered -- `torch.distributed.all_gather` over whatever backend the
    process group was created with (NCCL on the GPU box, gloo in the CPU tests).  The math never touches a collective.

`bucket_by_shape` / `ShardedIssuer.verify_mixed` implement the streamed mixed-shape workload: items are bucketed by their
attribute-kind vector (one compiled shape program per bucket), cut into chunks of at most `max_batch`, verified chunk by
chunk and scattered back into the caller's order.
"""
import numpy as np

from .issuer import PresentationBatch


def slice_bounds(total: int, rank: int, world: int):
    """Contiguous slice of `total` items owned by `rank`: [rank*total/world, (rank+1)*total/world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * total) // world, ((rank + 1) * total) // world


def pack_bitmap(verdicts) -> np.ndarray:
    """verdict bytes (0 = Ok, non-zero = VerificationFailure) -> little-endian bitmap, 1 bit per item (1 = rejected)."""
    return np.packbits(np.asarray(verdicts, dtype=np.uint8) != 0, bitorder="little")


def unpack_bitmap(bitmap, count: int) -> np.ndarray:
    return np.unpackbits(np.asarray(bitmap, dtype=np.uint8), count=count, bitorder="little").astype(np.uint8)


def bucket_by_shape(kinds_list):
    """kinds_list: one attribute-kind vector (bytes) per item -> {kinds: ndarray of item indices, in stream order}."""
    buckets = {}
    for i, k in enumerate(kinds_list):
        buckets.setdefault(bytes(k), []).append(i)
    return {k: np.asarray(v, dtype=np.int64) for k, v in buckets.items()}


class ShardedIssuer:
    """Data-parallel wrapper around one `Issuer` per process.  With no process group (world 1) it degenerates to the
    single-GPU path, so the same code drives 1, 2, 4 or 8 B200s."""

    def __init__(self, issuer, rank=None, world=None, group=None):
        self.issuer = issuer
        self.group = group
        if rank is None or world is None:
            rank, world = 0, 1
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank, world = dist.get_rank(group), dist.get_world_size(group)
            except ImportError:
                pass
        self.rank, self.world = rank, world

    # -- the only communication on the path: the verdict bitmap ------------------------------------------------------
    def _gather_bitmaps(self, local_bits: np.ndarray, total: int) -> np.ndarray:
        if self.world == 1:
            return unpack_bitmap(local_bits, total)
        import torch
        import torch.distributed as dist
        width = (-(-total // self.world) + 7) // 8 + 1          # bytes of the largest slice's bitmap
        dev = "cuda" if dist.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.zeros(width, dtype=torch.uint8, device=dev)
        mine[:len(local_bits)] = torch.from_numpy(local_bits).to(dev)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        out = np.empty(total, np.uint8)
        for r, p in enumerate(parts):
            lo, hi = slice_bounds(total, r, self.world)
            out[lo:hi] = unpack_bitmap(p.cpu().numpy(), hi - lo)
        return out

    def local_slice(self, batch: PresentationBatch) -> PresentationBatch:
        lo, hi = slice_bounds(batch.count, self.rank, self.world)
        return PresentationBatch(batch.kinds, batch.fields[:, lo:hi])

    def verify_batch(self, batch: PresentationBatch) -> np.ndarray:
        """Batch Issuer::verify over the global batch: this rank verifies its slice; every rank returns all verdicts."""
        local = self.issuer.verify_batch(self.local_slice(batch)) if batch.count else np.zeros(0, np.uint8)
        return self._gather_bitmaps(pack_bitmap(local), batch.count)

    def verify_issuance_batch(self, batch: PresentationBatch) -> np.ndarray:
        local = self.issuer.verify_issuance_batch(self.local_slice(batch)) if batch.count else np.zeros(0, np.uint8)
        return self._gather_bitmaps(pack_bitmap(local), batch.count)

    def verify_mixed(self, kinds_list, items, issuers=None) -> np.ndarray:
        """Streamed verification of presentations of mixed shapes (BASELINE config 5).

        kinds_list[i] is item i's attribute-kind vector and items[i] its words, uint8 [n_fields(kinds)][32].  Shapes with a
        different attribute count belong to different issuers: pass `issuers` = {n_attrs: Issuer}; by default every item
        goes to self.issuer.  Items are bucketed by shape, verified in chunks of at most max_batch, and the verdicts are
        returned in stream order."""
        total = len(kinds_list)
        out = np.zeros(total, np.uint8)
        for kinds, idx in sorted(bucket_by_shape(kinds_list).items()):
            issuer = issuers[len(kinds)] if issuers else self.issuer
            step = issuer.max_batch * self.world
            for s in range(0, len(idx), step):
                sel = idx[s:s + step]
                fields = np.ascontiguousarray(np.stack([items[i] for i in sel], axis=1))
                lo, hi = slice_bounds(len(sel), self.rank, self.world)
                local = issuer.verify_batch(PresentationBatch(kinds, fields[:, lo:hi])) if hi > lo else np.zeros(0, np.uint8)
                out[sel] = self._gather_bitmaps(pack_bitmap(local), len(sel))
        return out


class MultiGpuIssuer:
    """ONE process driving several B200s: a replicated `Issuer` context per device and one host thread per context (SURVEY 8e:
    "one thread + stream set per GPU").  This is the form a single-process caller -- the reference is a library, its
    `Issuer::verify_batch` shim runs inside the application -- uses instead of one process per GPU; the partition is the same
    (contiguous item slices, no data-path communication, verdicts concatenated).  The C ABI is thread-compatible across contexts:
    each context owns its streams, workspace and shape cache, and every call selects its device first; ctypes releases the GIL
    for the duration of a call, so the device passes run concurrently."""

    def __init__(self, system_parameters, issuer_parameters, amacs_key=None, devices=(0,), max_batch=65536, _binding=None):
        from concurrent.futures import ThreadPoolExecutor
        from .issuer import Issuer
        if not devices:
            raise ValueError("at least one device")
        self.devices = list(devices)
        self.issuers = [Issuer(system_parameters, issuer_parameters, amacs_key, device=d, max_batch=max_batch, _binding=_binding) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.devices))

    def host_array(self, shape):
        return self.issuers[0].host_array(shape)

    def _fan_out(self, method, batch):
        g = len(self.issuers)
        if batch.count == 0:
            return np.zeros(0, np.uint8)
        jobs = []
        for k, iss in enumerate(self.issuers):
            lo, hi = slice_bounds(batch.count, k, g)
            if hi > lo:
                jobs.append(self._pool.submit(getattr(iss, method), PresentationBatch(batch.kinds, batch.fields[:, lo:hi])))
        return np.concatenate([j.result() for j in jobs])

    def verify_batch(self, batch: PresentationBatch) -> np.ndarray:
        """Batch Issuer::verify: device k verifies items [k*N/G, (k+1)*N/G); verdicts in item order."""
        return self._fan_out("verify_batch", batch)

    def verify_issuance_batch(self, batch: PresentationBatch) -> np.ndarray:
        return self._fan_out("verify_issuance_batch", batch)

    def close(self):
        self._pool.shutdown(wait=True)
        for iss in self.issuers:
            iss.close()
