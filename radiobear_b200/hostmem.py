"""Page-locked host buffers for results and large inputs (torch pinned memory as plumbing)."""
import numpy as np


class PinnedPool:
    """Page-locked host result buffers (torch pinned tensors when torch + a GPU are there).

    A buffer is handed out as a numpy view and re-used only after the caller dropped every view of it
    (every view keeps a reference to the root array, so its refcount tells), so results of earlier runs
    are never overwritten behind the user's back, while steady-state loops pay neither a 92 MB
    allocation nor a pageable-memory D2H."""

    def __init__(self, pin=True):
        self._roots = []     # root uint8 arrays (each keeps its pinned tensor alive)
        self._pin = pin

    def _alloc(self, nbytes):
        if self._pin:
            try:
                import torch
                return torch.empty(nbytes, dtype=torch.uint8, pin_memory=True).numpy()
            except (ImportError, RuntimeError):
                pass
        return np.empty(nbytes, dtype=np.uint8)

    def get(self, shape, dtype):
        import sys
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if nbytes < (1 << 20):
            return np.empty(shape, dtype=dtype)
        root = None
        for r in self._roots:
            # refs: the list, the loop variable, getrefcount's argument -> 3 means nobody else holds a view
            if r.size >= nbytes and sys.getrefcount(r) <= 3:
                root = r
                break
        if root is None:
            self._roots = [r for r in self._roots if sys.getrefcount(r) > 3][-3:]
            root = self._alloc(nbytes)
            self._roots.append(root)
        return root[:nbytes].view(dtype).reshape(shape)


pinned_pool = PinnedPool()


def pinned_copy(a):
    """Private page-locked copy of a (large) array: H2D copies from it run at full PCIe speed."""
    out = PinnedPool()._alloc(a.nbytes)[:a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out
