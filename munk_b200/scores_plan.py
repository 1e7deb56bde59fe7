"""Host-side integer planning for separate_scores (reference: munk/munk/munk.py:20-68).

The reference selects with boolean masks; all five outputs are row-major traversals of
structured cell sets.  From the two index-pair lists this module derives, exactly and with
integer arithmetic only, where every selected cell lands, so that the device can do the
selection in one streaming pass (munk_separate_scores in munk_b200/csrc/scores.cu).
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib


def _csr(n_rows, rows, cols):
    """CSR (ptr, cols) of the cell list, rows ascending and columns ascending inside a row."""
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(ptr, rows + 1, 1)
    return np.cumsum(ptr), rows, cols


class SeparatePlan:
    def __init__(self, n1, n2, landmark_pair_idxs, homolog_pair_idxs):
        ls, lt = (np.asarray(x, dtype=np.int64) for x in zip(*landmark_pair_idxs))
        hs, ht = (np.asarray(x, dtype=np.int64) for x in zip(*homolog_pair_idxs))
        # numpy fancy indexing accepts negative indices; normalise like numpy does
        ls, hs = np.where(ls < 0, ls + n1, ls), np.where(hs < 0, hs + n1, hs)
        lt, ht = np.where(lt < 0, lt + n2, lt), np.where(ht < 0, ht + n2, ht)
        for arr, bound in ((ls, n1), (hs, n1), (lt, n2), (ht, n2)):
            if arr.size and (arr.min() < 0 or arr.max() >= bound):
                raise IndexError("index out of bounds for scores of shape (%d, %d)" % (n1, n2))
        self.n1, self.n2 = n1, n2
        self.ls, self.lt = ls, lt

        rowL = np.zeros(n1, dtype=np.uint8); rowL[ls] = 1
        colL = np.zeros(n2, dtype=np.uint8); colL[lt] = 1
        rL, cL = int(rowL.sum()), int(colL.sum())
        colrank = np.zeros(n2, dtype=np.int64)
        colrank[1:] = np.cumsum(colL[:-1])

        # landmark-pair cells (a set) and H_H cells = homolog cells minus landmark-pair cells
        pair_keys = np.unique(ls * n2 + lt)
        hom_keys = np.setdiff1d(np.unique(hs * n2 + ht), pair_keys, assume_unique=True)
        p_ptr, _, p_col = _csr(n1, pair_keys // n2, pair_keys % n2)
        h_ptr, h_row, h_col = _csr(n1, hom_keys // n2, hom_keys % n2)

        # per-row element counts of the three streamed outputs
        pairs_in_row = np.diff(p_ptr)
        h_in_other = np.zeros(n1, dtype=np.int64)
        sel = (rowL[h_row] == 0) & (colL[h_col] == 0)
        np.add.at(h_in_other, h_row[sel], 1)
        is_l = rowL.astype(bool)
        cnt_lloff = np.where(is_l, cL - pairs_in_row, 0)
        cnt_lnonl = np.where(is_l, n2 - cL, cL)
        cnt_other = np.where(is_l, 0, (n2 - cL) - h_in_other)

        def offsets(cnt):
            off = np.zeros(n1, dtype=np.int64)
            off[1:] = np.cumsum(cnt[:-1])
            return off, int(cnt.sum())

        self.off_lloff, self.n_lloff = offsets(cnt_lloff)
        self.off_lnonl, self.n_lnonl = offsets(cnt_lnonl)
        self.off_other, self.n_other = offsets(cnt_other)
        self.rowL, self.colL, self.colrank = rowL, colL, colrank.astype(np.int32)
        self.p_ptr, self.p_col = p_ptr.astype(np.int32), p_col.astype(np.int32)
        self.h_ptr, self.h_col = h_ptr.astype(np.int32), h_col.astype(np.int32)
        self.h_row = h_row.astype(np.int32)
        self.n_hh = int(hom_keys.size)
        assert self.n_lloff == rL * cL - pair_keys.size
        assert self.n_lnonl == rL * (n2 - cL) + (n1 - rL) * cL

    # ---- device side ----------------------------------------------------------------------
    def _dev(self, dev, arr):
        """Device copy of a plan array, uploaded once per (plan, device)."""
        cache = self.__dict__.setdefault("_dev_cache", {})
        key = (dev.index, id(arr))
        t = cache.get(key)
        if t is None:
            t = cache[key] = (arr, torch.from_numpy(np.ascontiguousarray(arr)).to(dev.torch_device))
        return t[1]

    def run(self, dev, S, to_host=True):
        """Five arrays as separate_scores returns them (numpy if to_host else CUDA tensors)."""
        d = lambda a: self._dev(dev, a)  # noqa: E731
        f64 = dict(dtype=torch.float64, device=dev.torch_device)
        ll_diag = dev.gather_pairs(S, self.ls.astype(np.int32), self.lt.astype(np.int32))
        hh = dev.gather_pairs(S, self.h_row, self.h_col) if self.n_hh else torch.empty(0, **f64)
        o_lloff = torch.empty(self.n_lloff, **f64)
        o_lnonl = torch.empty(self.n_lnonl, **f64)
        o_other = torch.empty(self.n_other, **f64)
        bufs = [d(self.rowL), d(self.colL), d(self.colrank), d(self.p_ptr), d(self.p_col),
                d(self.h_ptr), d(self.h_col), d(self.off_lloff), d(self.off_lnonl), d(self.off_other)]
        p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        check(lib.munk_separate_scores(self.n1, self.n2, p(S), S.stride(0), *[p(b) for b in bufs],
                                       p(o_lloff), p(o_lnonl), p(o_other), dev.stream()),
              "munk_separate_scores")
        outs = (ll_diag, o_lloff, o_lnonl, hh, o_other)
        if not to_host:
            return outs
        torch.cuda.current_stream(dev.index).synchronize()
        return tuple(t.cpu().numpy() for t in outs)

    def means(self, dev, S):
        """(mean(H_H) - mean(other), mean(other), mean(H_H)) as permtest.py:58-66 returns them;
        computed on the device without materialising the selections.  Returns a CUDA tensor of
        5 doubles {hom_mean, other_mean, diff, |H_H|, |other|} (no synchronisation)."""
        d = lambda a: self._dev(dev, a)  # noqa: E731
        p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        rowL, colL, hr, hc = d(self.rowL), d(self.colL), d(self.h_row), d(self.h_col)
        scratch = torch.empty(self.n1, dtype=torch.float64, device=dev.torch_device)
        out = torch.empty(5, dtype=torch.float64, device=dev.torch_device)
        check(lib.munk_separate_scores_means(self.n1, self.n2, p(S), S.stride(0), p(rowL), p(colL),
                                             self.n_hh, p(hr), p(hc), float(self.n_other), p(scratch),
                                             p(out), dev.stream()), "munk_separate_scores_means")
        return out


def plan_separate_scores(n1, n2, landmark_pair_idxs, homolog_pair_idxs):
    return SeparatePlan(n1, n2, landmark_pair_idxs, homolog_pair_idxs)
