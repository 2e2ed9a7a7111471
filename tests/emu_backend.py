"""CPU emulator of the kernel layer — TEST INFRASTRUCTURE ONLY.

Implements the `flox_b200.kernels.CudaKernels` interface with numpy on CPU torch tensors so that
the package's HOST logic (dtype / fill rules, chunk driver, map-reduce plumbing, gloo world-size-2
paths) can be tested on a machine without a GPU.  It restates what each `floxb_*` entry point of
include/flox_b200.h promises; it is installed with `flox_b200.kernels.set_backend` by the CPU test
fixtures and is never imported by the package.  GPU tests (`-m gpu`) never use it.
"""
from __future__ import annotations

import numpy as np
import torch

F32, F64, I32, I64 = 0, 1, 2, 3
INT64_MAX = np.iinfo(np.int64).max
INT64_MIN = np.iinfo(np.int64).min


def _np(t):
    return None if t is None else t.numpy()


def _key_enc(v: np.ndarray) -> np.ndarray:
    if v.dtype == np.float32:
        b = v.view(np.int32)
        return (b ^ ((b >> 31) & 0x7FFFFFFF)).astype(np.int64)
    if v.dtype == np.float64:
        b = v.view(np.int64)
        return b ^ ((b >> 63) & 0x7FFFFFFFFFFFFFFF)
    return v.astype(np.int64)


def _key_dec(k: np.ndarray, dtype) -> np.ndarray:
    if dtype == np.float32:
        b = k.astype(np.int32)
        return (b ^ ((b >> 31) & 0x7FFFFFFF)).view(np.float32)
    if dtype == np.float64:
        return (k ^ ((k >> 63) & 0x7FFFFFFFFFFFFFFF)).view(np.float64)
    return k.astype(dtype)


class EmuKernels:
    name = "emu"

    def __init__(self):
        self.launches = 0

    def _iter(self, chunks, A, B, G):
        """Yield (values[A,n,B] float64-or-int64 view, valid mask, bins[A,n,B]) per chunk."""
        for v, c in chunks:
            n = c.numel()
            vals = v.numpy().reshape(A, n, B)
            codes = c.numpy().astype(np.int64)
            ok = (codes >= 0) & (codes < G)
            ab = np.arange(A)[:, None, None] * B + np.arange(B)[None, None, :]
            if B == 1:
                bins = ab * G + np.where(ok, codes, 0)[None, :, None]
            else:  # group-major [G][A*B] (include/flox_b200.h)
                bins = np.where(ok, codes, 0)[None, :, None] * (A * B) + ab
            valid = np.broadcast_to(ok[None, :, None], vals.shape)
            yield vals, valid, bins

    def moments(self, chunks, A, B, G, op, acc, cnt, nnan, init=True, force_path=0):
        self.launches += 1
        a, c, m = _np(acc), _np(cnt), _np(nnan)
        if init:
            a[...] = 1 if op == 1 else 0
            c[...] = 0
            m[...] = 0
        for vals, valid, bins in self._iter(chunks, A, B, G):
            isn = np.isnan(vals) if vals.dtype.kind == "f" else np.zeros(vals.shape, bool)
            good = valid & ~isn
            w = vals[good].astype(a.dtype)
            if op == 2:
                w = w * w
            if op == 1:
                np.multiply.at(a, bins[good], w)
            else:
                np.add.at(a, bins[good], w)
            np.add.at(c, bins[good], 1)
            np.add.at(m, bins[valid & isn], 1)

    def minmax(self, chunks, A, B, G, is_max, key, cnt, nnan, init=True, force_path=0, value_dtype=None):
        self.launches += 1
        k, c, m = _np(key), _np(cnt), _np(nnan)
        if init:
            k[...] = INT64_MIN if is_max else INT64_MAX
            c[...] = 0
            m[...] = 0
        for vals, valid, bins in self._iter(chunks, A, B, G):
            isn = np.isnan(vals) if vals.dtype.kind == "f" else np.zeros(vals.shape, bool)
            good = valid & ~isn
            keys = _key_enc(np.ascontiguousarray(vals[good]))
            (np.maximum if is_max else np.minimum).at(k, bins[good], keys)
            np.add.at(c, bins[good], 1)
            np.add.at(m, bins[valid & isn], 1)

    def group_index(self, codes, G, perm, goff, cursor):
        self.launches += 1
        c = codes.numpy().astype(np.int64)
        ok = (c >= 0) & (c < G)
        rows = np.nonzero(ok)[0]
        order = rows[np.argsort(c[rows], kind="stable")]
        p = _np(perm)
        p[: order.size] = order
        counts = np.bincount(c[rows], minlength=G) if G else np.zeros(0, dtype=np.int64)
        _np(goff)[...] = np.concatenate([[0], np.cumsum(counts)])

    def reduce_cols(self, values, perm, goff, A, n, B, G, mode, acc, cnt, nnan, init=True, group_major=False):
        self.launches += 1
        v = values.numpy().reshape(A, n, B)
        p, off = _np(perm), _np(goff)
        if group_major:  # [G][A*B] storage: expose it as [A, B, G] views
            a_, c_, m_ = (_np(t).reshape(G, A, B).transpose(1, 2, 0) for t in (acc, cnt, nnan))
        else:
            a_, c_, m_ = _np(acc).reshape(A, B, G), _np(cnt).reshape(A, B, G), _np(nnan).reshape(A, B, G)
        ident = {0: 0, 1: 1, 2: 0, 3: INT64_MIN, 4: INT64_MAX}[mode]
        for g in range(G):
            rows = p[off[g]:off[g + 1]]
            x = v[:, rows, :]
            isn = np.isnan(x) if x.dtype.kind == "f" else np.zeros(x.shape, bool)
            good = ~isn
            if mode in (0, 1, 2):
                w = x.astype(a_.dtype)
                if mode == 2:
                    w = w * w
                w = np.where(good, w, ident)
                r = w.prod(axis=1) if mode == 1 else w.sum(axis=1)
            else:
                keys = _key_enc(np.ascontiguousarray(np.where(isn, 0, x).astype(x.dtype))).reshape(x.shape)
                keys = np.where(good, keys, ident)
                r = keys.max(axis=1, initial=ident) if mode == 3 else keys.min(axis=1, initial=ident)
            cg, mg = good.sum(axis=1), isn.sum(axis=1)
            if init:
                a_[:, :, g], c_[:, :, g], m_[:, :, g] = r, cg, mg
            else:
                if mode in (0, 2):
                    a_[:, :, g] += r
                elif mode == 1:
                    a_[:, :, g] *= r
                elif mode == 3:
                    a_[:, :, g] = np.maximum(a_[:, :, g], r)
                else:
                    a_[:, :, g] = np.minimum(a_[:, :, g], r)
                c_[:, :, g] += cg
                m_[:, :, g] += mg

    def group_means(self, value_dtype, acc, cnt, nnan, skipna, means):
        self.launches += 1
        a, c, m, out = _np(acc), _np(cnt), _np(nnan), _np(means)
        with np.errstate(invalid="ignore", divide="ignore"):
            if value_dtype == torch.float32:
                mu = (a.astype(np.float32).astype(np.float64) / c).astype(np.float32).astype(np.float64)
            else:
                mu = a.astype(np.float64) / c
        if not skipna:
            mu = np.where(m > 0, np.nan, mu)
        out[...] = mu

    def sqdev(self, chunks, A, B, G, means, skipna, m2, init=True, value_dtype=None):
        self.launches += 1
        mu, out = _np(means), _np(m2)
        if init:
            out[...] = 0
        for vals, valid, bins in self._iter(chunks, A, B, G):
            isn = np.isnan(vals) if vals.dtype.kind == "f" else np.zeros(vals.shape, bool)
            sel = valid & ~isn if skipna else valid
            d = vals[sel].astype(np.float64) - mu[bins[sel]]
            np.add.at(out, bins[sel], d * d)

    def argext(self, chunks, A, B, G, key, nnan, skipna, index_base, idx, init=True, value_dtype=None):
        self.launches += 1
        k, m, out = _np(key), _np(nnan), _np(idx)
        if init:
            out[...] = INT64_MAX
        for (vals, valid, bins), base in zip(self._iter(chunks, A, B, G), index_base):
            n = vals.shape[1]
            pos = np.broadcast_to((base + np.arange(n))[None, :, None], vals.shape)
            isn = np.isnan(vals) if vals.dtype.kind == "f" else np.zeros(vals.shape, bool)
            keys = _key_enc(np.ascontiguousarray(np.where(isn, 0, vals).astype(vals.dtype)))
            poisoned = (m[bins] > 0) & (not skipna)
            hit = valid & np.where(poisoned, isn, (~isn) & (keys == k[bins]))
            np.minimum.at(out, bins[hit], pos[hit])

    def finalize(self, kind, value_dtype, acc, acc2, cnt, nnan, skipna, fill, min_count, mc_fill, ddof, out, group_major_ab=0):
        self.launches += 1
        o = _np(out)
        if group_major_ab:  # accumulators are [G][AB]; results go out as [AB][G]
            gdim = o.size // group_major_ab

            class _T:  # transposed read-only views with the torch-ish .numpy() the helpers expect
                def __init__(self, t):
                    self.a = None if t is None else np.ascontiguousarray(t.numpy().reshape(gdim, group_major_ab).T).reshape(-1)

                def numpy(self):
                    return self.a

            acc, acc2, cnt, nnan = (None if t is None else _T(t) for t in (acc, acc2, cnt, nnan))
        c = _np(cnt).astype(np.int64)
        m = _np(nnan).astype(np.int64) if nnan is not None else np.zeros_like(c)
        rows = c + m
        odt = o.dtype
        is_f = odt.kind == "f"

        def conv(v):
            if v is None:
                return np.array(np.nan if is_f else 0).astype(odt)
            if not is_f:
                try:
                    return np.array(int(v)).astype(odt)
                except (ValueError, OverflowError):
                    return np.array(0, dtype=odt)
            return np.array(float(v)).astype(odt)

        fillv, mcv = conv(fill), conv(mc_fill)
        nan = np.array(np.nan).astype(odt) if is_f else np.array(0, dtype=odt)
        poisoned = (m > 0) & (not skipna)
        with np.errstate(all="ignore"):
            if kind == 0:  # SUM
                r = np.where(rows == 0, fillv, np.where(poisoned, nan, _np(acc).astype(odt)))
            elif kind == 1:  # NANLEN
                r = np.where(rows == 0, fillv, c.astype(odt))
            elif kind == 2:  # MEAN
                a = _np(acc)
                s = a.astype(np.float32).astype(np.float64) if odt == np.float32 else a.astype(np.float64)
                s = np.where(rows == 0, np.float64(fillv), np.where(poisoned, np.nan, s))
                r = (s / c.astype(np.float64)).astype(odt)
            elif kind == 8:  # PICK
                r = np.where((rows == 0) | ((c == 0) if skipna else False), fillv, _np(acc).astype(odt))
            elif kind in (3, 4):  # MAX / MIN
                vdt = {torch.float32: np.float32, torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[value_dtype]
                v = _key_dec(_np(acc), vdt)
                r = v.astype(odt)
                if np.dtype(vdt).kind == "f" and skipna:
                    ident = -np.inf if kind == 3 else np.inf
                    r = np.where(v == ident, fillv, r)
                r = np.where(c == 0, fillv, r)
                r = np.where(poisoned, nan, r)
                r = np.where(rows == 0, fillv, r)
            elif kind in (5, 6):  # VAR / STD
                den = (c - ddof).astype(np.float64)
                v = np.where(poisoned, np.nan, _np(acc) / den)
                v = np.where(den < 0, np.nan, v)
                if kind == 6:
                    v = np.sqrt(v)
                r = v.astype(odt)
            elif kind == 7:  # ARG
                ix = _np(acc)
                r = np.where(ix == INT64_MAX, fillv, ix.astype(odt))
            else:
                raise ValueError(kind)
        if min_count > 0:
            r = np.where(c < min_count, mcv, r)
        o[...] = r.astype(odt)

    def combine(self, kind, stacked, out):
        self.launches += 1
        s = _np(stacked)
        with np.errstate(all="ignore"):
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fn = {0: np.sum, 1: np.max, 2: np.min, 3: np.prod, 4: np.nanmax, 5: np.nanmin}[kind]
                _np(out)[...] = fn(s, axis=0)

    def var_combine(self, m2, s, n, out_m2, out_s, out_n):
        self.launches += 1
        M2, S, N = _np(m2), _np(s), _np(n)
        m, ss, nn = M2[0].copy(), S[0].copy(), N[0].copy()
        adj = np.zeros_like(m)
        with np.errstate(all="ignore"):
            for k in range(1, M2.shape[0]):
                deg = (nn == 0) | (N[k] == 0)
                num = nn * S[k] - N[k] * ss
                adj = adj + num * num / (nn * N[k] * (nn + N[k]) + deg.astype(np.float64))
                m, ss, nn = m + M2[k], ss + S[k], nn + N[k]
        _np(out_m2)[...] = m + adj
        _np(out_s)[...] = ss
        _np(out_n)[...] = nn

    def digitize(self, labels, edges, closed_right, codes):
        self.launches += 1
        x = _np(labels).astype(np.float64)
        e = _np(edges)
        if e.size < 2:
            _np(codes)[...] = -1
            return
        idx = np.digitize(x, e, right=bool(closed_right)) - 1
        inside = (x > e[0]) & (x <= e[-1]) if closed_right else (x >= e[0]) & (x < e[-1])
        _np(codes)[...] = np.where(inside, idx, -1)

    def ravel_codes(self, c0, c1, size0, size1, out):
        self.launches += 1
        a, b = _np(c0), _np(c1)
        bad = (a < 0) | (b < 0) | (a >= size0) | (b >= size1)
        _np(out)[...] = np.where(bad, -1, a * size1 + b)

    def label_minmax(self, labels):
        self.launches += 1
        a = _np(labels)
        return int(a.min()), int(a.max())

    def label_presence(self, labels, lo, span):
        self.launches += 1
        rel = _np(labels).astype(np.int64) - lo
        table = np.zeros(span, dtype=np.uint8)
        table[rel[(rel >= 0) & (rel < span)]] = 1
        return torch.from_numpy(table)

    def lut_codes(self, labels, lo, lut, codes):
        self.launches += 1
        rel = _np(labels).astype(np.int64) - lo
        inside = (rel >= 0) & (rel < lut.numel())
        _np(codes)[...] = np.where(inside, _np(lut)[np.clip(rel, 0, max(lut.numel() - 1, 0))] if lut.numel() else -1, -1)

    def smem_group_limit(self, value_dtype, family):
        return 14000

    def max_chunks_per_launch(self):
        return 128

    @staticmethod
    def host_pack_codes(codes_np, n_groups, out):
        _np(out)[...] = np.where((codes_np >= 0) & (codes_np < n_groups), codes_np, -1)

    def widen_codes(self, codes16):
        return codes16.to(torch.int32)

    def firstlast(self, chunks, A, B, G, is_last, skipna, index_base, pos, init=True, value_dtype=None):
        self.launches += 1
        p = _np(pos)
        if init:
            p[...] = INT64_MIN if is_last else INT64_MAX
        for (vals, valid, bins), base in zip(self._iter(chunks, A, B, G), index_base):
            n = vals.shape[1]
            idx = np.broadcast_to((base + np.arange(n, dtype=np.int64))[None, :, None], vals.shape)
            ok = valid & ~(np.isnan(vals) if (skipna and vals.dtype.kind == "f") else np.zeros(vals.shape, bool))
            (np.maximum if is_last else np.minimum).at(p, bins[ok], idx[ok])

    def take(self, chunks, A, B, G, pos, index_base, acc, init=True, value_dtype=None):
        self.launches += 1
        a, p = _np(acc), _np(pos)
        if init:
            a[...] = 0
        for (vals, valid, bins), base in zip(self._iter(chunks, A, B, G), index_base):
            n = vals.shape[1]
            idx = np.broadcast_to((base + np.arange(n, dtype=np.int64))[None, :, None], vals.shape)
            hit = valid & (p[bins] == idx)
            a[bins[hit]] = vals[hit].astype(a.dtype)

    # -- sorted-segment family (csrc/segments.cu) ----------------------------------------------------
    def segment_valid(self, sorted_vals, starts, nvalid):
        self.launches += 1
        v, s = _np(sorted_vals), _np(starts)
        out = _np(nvalid).reshape(v.shape[0], len(s) - 1)
        for g in range(len(s) - 1):
            out[:, g] = (~np.isnan(v[:, s[g]:s[g + 1]])).sum(axis=1)

    def segment_quantile(self, sorted_vals, starts, nvalid, before, q, skipna, fill, out):
        self.launches += 1
        v, s, qs = _np(sorted_vals), _np(starts), _np(q)
        R, G = v.shape[0], len(s) - 1
        nv, bef = _np(nvalid).reshape(R, G), _np(before).reshape(R, G)
        o = _np(out).reshape(len(qs), R, G)
        rows = np.arange(R)
        with np.errstate(all="ignore"):
            for g in range(G):
                ln = s[g + 1] - s[g]
                for iq, qq in enumerate(qs):
                    if ln == 0:
                        o[iq, :, g] = fill
                        continue
                    virt = qq * (nv[:, g] - 1).astype(np.float64)
                    virt = np.where(bef[:, g] != 0, virt + bef[:, g].astype(np.float64), virt)
                    fl, ce = np.floor(virt), np.ceil(virt)
                    gamma = virt - fl
                    li = np.clip(fl.astype(np.int64) - bef[:, g], 0, ln - 1) + s[g]
                    hi = np.clip(ce.astype(np.int64) - bef[:, g], 0, ln - 1) + s[g]
                    a, b = v[rows, li], v[rows, hi]
                    d = (b - a).astype(v.dtype)
                    res = a.astype(np.float64) + d.astype(np.float64) * gamma
                    res = np.where(gamma >= 0.5, b.astype(np.float64) - d.astype(np.float64) * (1.0 - gamma), res)
                    bad = nv[:, g] == 0
                    if not skipna:
                        bad = bad | (nv[:, g] != ln)
                    o[iq, :, g] = np.where(bad, np.nan, res).astype(o.dtype)

    def segment_scan(self, kind, values, perm, sorted_codes, skipna, out):
        """A true segmented scan in f64 / i64 per group (what csrc/segments.cu does), rounded once."""
        self.launches += 1
        v, p, sc = _np(values), _np(perm), _np(sorted_codes)
        o = _np(out)
        n = v.shape[1]
        if n == 0:
            return
        heads = np.flatnonzero(np.concatenate(([True], sc[1:] != sc[:-1])))
        bounds = np.concatenate((heads, [n]))
        sv = v[:, p]
        res = np.empty_like(sv)
        isf = sv.dtype.kind == "f"
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            seg = sv[:, lo:hi]
            if kind == 0:
                wide = seg.astype(np.float64 if isf else np.int64)
                if skipna and isf:
                    wide = np.where(np.isnan(wide), 0.0, wide)
                res[:, lo:hi] = np.cumsum(wide, axis=1).astype(sv.dtype)
            else:
                if isf:
                    idx = np.where(np.isnan(seg), 0, np.arange(hi - lo))
                    np.maximum.accumulate(idx, axis=1, out=idx)
                    res[:, lo:hi] = np.take_along_axis(seg, idx, axis=1)
                else:
                    res[:, lo:hi] = seg
        o[:, p] = res
