"""ctypes mirror of include/zmo_b200.h (one class per opaque handle, numpy in / numpy out)."""
import ctypes as C
import os
import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)


def lib_path():
    return os.path.join(PKG_DIR, "lib", "libzmo_b200.so")


def wtzmo_path():
    return os.path.join(PKG_DIR, "bin", "wtzmo")


class ZmoError(RuntimeError):
    pass


class ZmoParams(C.Structure):
    """zmo_params_t (include/zmo_b200.h); defaults are wtzmo's (wtzmo.c:1543-1588)."""
    _fields_ = [("hk", C.c_int32), ("hz", C.c_int32), ("ksize", C.c_int32), ("zsize", C.c_int32),
                ("ksave", C.c_int32), ("kovl", C.c_int32), ("zcut", C.c_int32), ("kvar", C.c_int32),
                ("kwin", C.c_int32), ("kstep", C.c_int32), ("zovl", C.c_int32), ("ztot", C.c_int32),
                ("w", C.c_int32), ("ew", C.c_int32), ("W", C.c_int32),
                ("M", C.c_int32), ("X", C.c_int32), ("O", C.c_int32), ("E", C.c_int32), ("T", C.c_int32),
                ("min_id", C.c_float),
                ("xvar", C.c_int32), ("yvar", C.c_int32), ("min_block_len", C.c_int32), ("max_overhang", C.c_int32),
                ("deviation_penalty", C.c_float), ("gap_penalty", C.c_float)]


def default_params(**kw):
    p = ZmoParams(hk=1, hz=1, ksize=16, zsize=10, ksave=4, kovl=300, zcut=64, kvar=2, kwin=800, kstep=400,
                  zovl=200, ztot=300, w=50, ew=800, W=3200, M=2, X=-5, O=-3, E=-1, T=-50, min_id=0.5,
                  xvar=128, yvar=64, min_block_len=160, max_overhang=256, deviation_penalty=1.0, gap_penalty=0.05)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    p.kstep = p.kwin // 2
    return p


class IndexStats(C.Structure):
    _fields_ = [("n_kmers", C.c_uint64), ("n_postings", C.c_uint64), ("n_filtered_high", C.c_uint64),
                ("n_indexed", C.c_uint64), ("kcut", C.c_uint32), ("kavg", C.c_uint32)]


DP_PROBLEM = np.dtype([("q_rid", "<u4"), ("t_rid", "<u4"),
                       ("q_start", "<i4"), ("q_step", "<i4"), ("q_comp", "<i4"), ("qlen", "<i4"),
                       ("t_start", "<i4"), ("t_step", "<i4"), ("t_comp", "<i4"), ("tlen", "<i4"),
                       ("init_score", "<i4"), ("W", "<i4")])
DP_RESULT = np.dtype([("score", "<i4"), ("qe", "<i4"), ("te", "<i4"), ("aln", "<i4"), ("mat", "<i4"), ("mis", "<i4"),
                      ("ins", "<i4"), ("del", "<i4"), ("cigar_off", "<u8"), ("n_cigar", "<u4"), ("_pad", "<u4"),
                      ("cells", "<u8")])
EVENT = np.dtype([("tkey", "<u4"), ("ol", "<u4")])
PAIR = np.dtype([("qid", "<u4"), ("cid", "<u4")])
PAIRSEED = np.dtype([("n_zpair", "<u4"), ("ovl", "<i4", (2,)), ("win_off", "<u4", (2,)), ("n_win", "<u4", (2,))])
WINDOW = np.dtype([("beg", "<i4", (2,)), ("end", "<i4", (2,))])
TASK = np.dtype([("pair_idx", "<u4"), ("dir", "<u4")])
RECORD = np.dtype([("ok", "<i4"), ("score", "<i4"), ("tb", "<i4"), ("te", "<i4"), ("qb", "<i4"), ("qe", "<i4"),
                   ("aln", "<i4"), ("mat", "<i4"), ("mis", "<i4"), ("ins", "<i4"), ("del", "<i4"), ("_pad", "<i4"),
                   ("cigar_off", "<u8"), ("n_cigar", "<u4"), ("_pad2", "<u4")])
DOTRES = np.dtype([("n_zpair", "<u4"), ("score", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("tb", "<i4"), ("te", "<i4"),
                   ("strand", "<i4")])

_lib = None


def load_lib(path=None):
    """Load libzmo_b200.so.  Raises ZmoError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or lib_path()
    if not os.path.exists(path):
        raise ZmoError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path)
    vp, u32p, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    lib.zmo_ctx_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(ZmoParams)]
    lib.zmo_ctx_destroy.argtypes = [vp]
    lib.zmo_ctx_destroy.restype = None
    lib.zmo_last_error.restype = C.c_char_p
    lib.zmo_kernel_launches.argtypes = [vp]
    lib.zmo_kernel_launches.restype = C.c_uint64
    lib.zmo_stage_ms.argtypes = [vp, C.POINTER(C.c_double)]
    lib.zmo_stage_ms.restype = None
    lib.zmo_counters.argtypes = [vp, u64p]
    lib.zmo_counters.restype = None
    lib.zmo_reads_upload.argtypes = [vp, vp, C.c_uint64, vp, vp, C.c_uint32]
    lib.zmo_index_build.argtypes = [vp, C.c_uint32, C.c_uint32, u32p, C.POINTER(IndexStats)]
    lib.zmo_candidates.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.zmo_pair_windows.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.zmo_pair_align.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.zmo_pair_align_text.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.zmo_ctx_clone.argtypes = [vp, C.POINTER(vp)]
    lib.zmo_set_refine.argtypes = [vp, C.c_int]
    lib.zmo_pair_dotmatrix.argtypes = [vp, vp, C.c_uint32, vp]
    lib.zmo_dp_extend.argtypes = [vp, C.c_int, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    lib.zmo_dp_global.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    if path == lib_path():
        _lib = lib
    return lib


def pack_reads(seqs):
    """Pack sequences (iterables of 0..3 codes or ACGT strings) into the reference BaseBank layout
    (dna.h:78,263): 32 bases per uint64, base i at bits ((~i)&31)*2 of word i>>5, reads concatenated."""
    arrs = []
    for s in seqs:
        if isinstance(s, (str, bytes)):
            b = np.frombuffer(s.encode() if isinstance(s, str) else s, dtype=np.uint8)
            lut = np.full(256, 0, np.uint8)
            for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
                lut[ch] = v
            arrs.append(lut[b])
        else:
            arrs.append(np.asarray(s, dtype=np.uint8))
    lens = np.array([len(a) for a in arrs], dtype=np.uint32)
    offs = np.zeros(len(arrs), dtype=np.uint64)
    if len(arrs) > 1:
        offs[1:] = np.cumsum(lens[:-1].astype(np.uint64))
    total = int(lens.astype(np.uint64).sum())
    allb = np.concatenate(arrs).astype(np.uint64) if arrs else np.zeros(0, np.uint64)
    nw = (total + 31) // 32 + 1
    pad = np.zeros(nw * 32, dtype=np.uint64)
    pad[:total] = allb
    shifts = ((31 - np.arange(32, dtype=np.uint64)) * 2).astype(np.uint64)
    words = np.bitwise_or.reduce(pad.reshape(nw, 32) << shifts, axis=1).astype(np.uint64)
    return words, total, offs, lens


class Zmo:
    """One device context (zmo_ctx).  All methods raise ZmoError on failure."""

    def __init__(self, params=None, device=0, lib=None, _clone_of=None):
        self.lib = lib or load_lib()
        self._h = C.c_void_p()
        if _clone_of is not None:
            self.params = _clone_of.params
            self._chk(self.lib.zmo_ctx_clone(_clone_of._h, C.byref(self._h)))
            self.n_reads, self.lens = _clone_of.n_reads, _clone_of.lens
            return
        self.params = params or default_params()
        self._chk(self.lib.zmo_ctx_create(C.byref(self._h), device, C.byref(self.params)))

    def clone(self):
        """zmo_ctx_clone: a context with its own streams / scratch that shares this one's reads and index."""
        return Zmo(lib=self.lib, _clone_of=self)

    def set_refine(self, on=True):
        """zmo_set_refine: -n, re-align stitched alignments with kswx_refine_alignment (set before cloning)."""
        self._chk(self.lib.zmo_set_refine(self._h, 1 if on else 0))

    def _chk(self, rc):
        if rc != 0:
            raise ZmoError("zmo error %d: %s" % (rc, self.lib.zmo_last_error().decode()))

    def close(self):
        if self._h:
            self.lib.zmo_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- bookkeeping
    def launches(self):
        return int(self.lib.zmo_kernel_launches(self._h))

    def stage_ms(self):
        a = (C.c_double * 12)()
        self.lib.zmo_stage_ms(self._h, a)
        return dict(zip(["index", "candidates", "pair_windows", "window_align", "gap_global", "end_extend", "dotmatrix", "copy", "dp_phase_wall"], list(a)))

    def counters(self):
        a = (C.c_uint64 * 8)()
        self.lib.zmo_counters(self._h, a)
        return dict(zip(["cells_ext", "cells_win", "cells_gap", "zpairs", "postings", "h2d_bytes", "d2h_bytes", "_"], [int(x) for x in a]))

    # -- stages
    def upload(self, words, n_bases, offs, lens):
        words = np.ascontiguousarray(words, np.uint64)
        offs = np.ascontiguousarray(offs, np.uint64)
        lens = np.ascontiguousarray(lens, np.uint32)
        self.n_reads = len(lens)
        self.lens = lens
        self._chk(self.lib.zmo_reads_upload(self._h, words.ctypes.data, int(n_bases), offs.ctypes.data, lens.ctypes.data, len(lens)))

    def upload_seqs(self, seqs):
        self.upload(*pack_reads(seqs))

    def index_build(self, beg=0, end=None, kcut=0):
        end = self.n_reads if end is None else end
        k = C.c_uint32(kcut)
        st = IndexStats()
        self._chk(self.lib.zmo_index_build(self._h, beg, end, C.byref(k), C.byref(st)))
        return int(k.value), st

    def candidates(self, qids):
        qids = np.ascontiguousarray(qids, np.uint32)
        off = np.zeros(len(qids) + 1, np.uint64)
        cap = max(1024, 256 * len(qids))
        while True:
            ev = np.zeros(cap, EVENT)
            need = C.c_uint64(0)
            rc = self.lib.zmo_candidates(self._h, qids.ctypes.data, len(qids), off.ctypes.data, ev.ctypes.data, cap, C.byref(need))
            if rc == -3:
                cap = int(need.value) + 16
                continue
            self._chk(rc)
            return off, ev[:int(need.value)]

    def pair_windows(self, pairs, slot=0):
        pairs = np.ascontiguousarray(pairs, PAIR)
        seeds = np.zeros(len(pairs), PAIRSEED)
        cap = max(1024, 32 * len(pairs))
        while True:
            wins = np.zeros(cap, WINDOW)
            need = C.c_uint64(0)
            rc = self.lib.zmo_pair_windows(self._h, slot, pairs.ctypes.data, len(pairs), seeds.ctypes.data, wins.ctypes.data, cap, C.byref(need))
            if rc == -3:
                cap = int(need.value) + 16
                continue
            self._chk(rc)
            return seeds, wins[:int(need.value)]

    def pair_align(self, tasks, slot=0, cigar_cap=1 << 20):
        tasks = np.ascontiguousarray(tasks, TASK)
        recs = np.zeros(len(tasks), RECORD)
        cap = cigar_cap
        while True:
            cig = np.zeros(cap, np.uint32)
            need = C.c_uint64(0)
            rc = self.lib.zmo_pair_align(self._h, slot, tasks.ctypes.data, len(tasks), recs.ctypes.data, cig.ctypes.data, cap, C.byref(need))
            if rc == -3:
                cap = int(need.value) + 16
                continue
            self._chk(rc)
            return recs, cig

    def pair_align_text(self, tasks, slot=0, text_cap=1 << 20):
        """zmo_pair_align_text: records whose cigar_off / n_cigar index the returned bytes (CIGAR text formatted on the device)."""
        tasks = np.ascontiguousarray(tasks, TASK)
        recs = np.zeros(len(tasks), RECORD)
        cap = text_cap
        while True:
            txt = np.zeros(cap, np.uint8)
            need = C.c_uint64(0)
            rc = self.lib.zmo_pair_align_text(self._h, slot, tasks.ctypes.data, len(tasks), recs.ctypes.data, txt.ctypes.data, cap, C.byref(need))
            if rc == -3:
                cap = int(need.value) + 16
                continue
            self._chk(rc)
            return recs, txt

    def pair_dotmatrix(self, pairs):
        pairs = np.ascontiguousarray(pairs, PAIR)
        out = np.zeros(len(pairs), DOTRES)
        self._chk(self.lib.zmo_pair_dotmatrix(self._h, pairs.ctypes.data, len(pairs), out.ctypes.data))
        return out

    def dp_extend(self, mode, probs):
        probs = np.ascontiguousarray(probs, DP_PROBLEM)
        res = np.zeros(len(probs), DP_RESULT)
        cap = int(np.maximum(probs["qlen"], 0).sum() + np.maximum(probs["tlen"], 0).sum() + 4 * len(probs) + 16)
        cig = np.zeros(cap, np.uint32)
        need = C.c_uint64(0)
        self._chk(self.lib.zmo_dp_extend(self._h, mode, probs.ctypes.data, len(probs), res.ctypes.data, cig.ctypes.data, cap, C.byref(need)))
        return res, cig

    def dp_global(self, probs, w):
        probs = np.ascontiguousarray(probs, DP_PROBLEM)
        w = np.ascontiguousarray(w, np.int32)
        res = np.zeros(len(probs), DP_RESULT)
        cap = int(np.maximum(probs["qlen"], 0).sum() + np.maximum(probs["tlen"], 0).sum() + 4 * len(probs) + 16)
        cig = np.zeros(cap, np.uint32)
        need = C.c_uint64(0)
        self._chk(self.lib.zmo_dp_global(self._h, probs.ctypes.data, w.ctypes.data, len(probs), res.ctypes.data, cig.ctypes.data, cap, C.byref(need)))
        return res, cig
