"""ctypes binding of libhl_b200.so (include/hlb200.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("HLB200_LIB", os.path.join(HERE, "libhl_b200.so"))   # HLB200_LIB: alternative build of the same library (tuning experiments)


class Hlb200Error(RuntimeError):
    pass


# ---- numpy dtypes mirroring the C structs of include/hlb200.h ----
MB_MOTION = np.dtype([("part_mode", "u1"), ("sub_mode", "u1", (4,)), ("ref_idx", "i1", (4,)), ("pad", "u1", (3,)), ("mv", "<i2", (4, 4, 2))])
MB_COEFFS = np.dtype([("luma_level", "<i2", (16, 16)), ("chroma_dc_level", "<i2", (2, 4)), ("chroma_ac_level", "<i2", (2, 4, 16)),
                      ("cbp_luma4x4", "<u2"), ("cbp_chroma_dc4x4", "u1", (2,)), ("cbp_chroma_ac4x4", "u1", (2,)), ("pad", "u1", (2,))])
SVC_STATE = np.dtype([("chroma_ac_level", "<i2", (2, 4, 16)), ("chroma_dc_level", "<i2", (2, 4))])   # hlb200_svc_mb_state_t
SVC_BASE_MB = np.dtype([("flags", "u1"), ("part_w", "u1"), ("part_h", "u1"), ("sub_w", "u1", (4,)), ("sub_h", "u1", (4,)), ("pred_flag", "i1", (4,)), ("ref_idx", "i1", (4,)),
                        ("pad", "u1"), ("mv", "<i2", (4, 4, 2))])   # hlb200_svc_base_mb_t
SVC_GEOM = np.dtype([("ref_width", "<i4"), ("ref_height", "<i4"), ("scaled_width", "<i4"), ("scaled_height", "<i4"), ("left_offset", "<i4"), ("top_offset", "<i4"),
                     ("level_idc", "<i4"), ("restricted", "<i4"), ("cropping_change", "<i4")])   # hlb200_svc_layer_geom_t
ME_CAND = np.dtype([("mb_x", "<i2"), ("mb_y", "<i2"), ("part_x", "u1"), ("part_y", "u1"), ("part_w", "u1"), ("part_h", "u1"), ("mv_x", "<i2"), ("mv_y", "<i2")])
ME_COST = np.dtype([("dist", "<i4"), ("bits_rest", "<i4"), ("single_ctr", "<i4"), ("cbp_luma4x4", "<u2"), ("total_coeff", "u1", (16,)),
                    ("trailing_ones", "u1", (16,)), ("pad", "<u2")])
MB_RECORD = np.dtype([
    ("mb_class", "u1"), ("mb_type", "u1"), ("part_mode", "u1"), ("sub_mode", "u1", (4,)), ("i16_pred_mode", "u1"), ("intra_chroma_pred_mode", "u1"),
    ("coded_block_pattern", "u1"), ("cbp_luma", "u1"), ("cbp_chroma", "u1"), ("cbp_chroma_dc4x4", "u1", (2,)), ("cbp_chroma_ac4x4", "u1", (2,)),
    ("cbp_luma4x4", "<u2"), ("mb_qp_delta", "i1"), ("qp_y", "u1"), ("qp_c", "u1", (2,)), ("ref_idx", "i1", (4,)), ("i4_pred_mode", "u1", (16,)),
    ("prev_intra4x4_pred_mode_flag", "u1", (16,)), ("rem_intra4x4_pred_mode", "u1", (16,)), ("pad", "u1", (3,)),
    ("mv", "<i2", (4, 4, 2)), ("mvd", "<i2", (4, 4, 2)), ("mad", "<i4"),
    ("luma_level", "<i2", (16, 16)), ("i16_dc_level", "<i2", (16,)), ("i16_ac_level", "<i2", (16, 16)),
    ("chroma_dc_level", "<i2", (2, 4)), ("chroma_ac_level", "<i2", (2, 4, 16)),
    ("me_trials", "<u4"), ("me_interp_ops", "<u4"), ("me_candidates", "<u2"), ("intra_trials", "<u2"), ("t_start_ns", "<u4"), ("t_end_ns", "<u4")], align=True)


class SliceParams(C.Structure):
    _fields_ = [("slice_type", C.c_int32), ("qp", C.c_int32), ("me_range", C.c_int32), ("num_refs", C.c_int32), ("chroma_qp_index_offset", C.c_int32),
                ("cur_slot", C.c_int32), ("ref_slot", C.c_int32 * 16), ("me_early_term_flag", C.c_int32), ("deblock_flag", C.c_int32)]


_lib = None


def load():
    """Loads libhl_b200.so; raises Hlb200Error when it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise Hlb200Error("libhl_b200.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; g.build()'`; there is no CPU fallback" % SO_PATH)
    lib = C.CDLL(SO_PATH)
    vp, ip = C.c_void_p, C.c_int
    sig = {
        "hlb200_init": [ip], "hlb200_device_count": [], "hlb200_version": [],
        "hlb200_stream_create": [ip, ip, ip, C.POINTER(vp)], "hlb200_stream_destroy": [vp], "hlb200_stream_set_cuda_stream": [vp, vp],
        "hlb200_stream_sync": [vp], "hlb200_frame_upload": [vp, vp, vp, vp, ip, ip], "hlb200_frame_set_device": [vp, vp, vp, vp], "hlb200_slot_upload": [vp, ip, vp, vp, vp],
        "hlb200_slot_download": [vp, ip, vp, vp, vp], "hlb200_state_reset": [vp],
        "hlb200_slice_encode": [vp, C.POINTER(SliceParams), vp], "hlb200_slice_encode_async": [vp, C.POINTER(SliceParams)], "hlb200_records_download": [vp, vp],
        "hlb200_slice_encode_batch_async": [C.POINTER(vp), C.POINTER(SliceParams), ip], "hlb200_slice_grid_size": [], "hlb200_slice_set_variant": [ip], "hlb200_slice_last_variant": [], "hlb200_slice_status": [vp, vp],
        "hlb200_interp_luma": [vp, ip, vp, vp], "hlb200_interp_chroma": [vp, ip, vp, vp, vp],
        "hlb200_tq_recon": [vp, ip, ip, vp, vp, vp, vp, vp, vp, vp], "hlb200_sad4x4": [vp, vp, ip, vp], "hlb200_homogeneity8x8": [vp, vp], "hlb200_dev_homogeneity8x8": [vp, ip, ip, vp, vp], "hlb200_me_cost": [vp, ip, ip, vp, ip, vp],
        "hlb200_dev_interp_luma": [vp, ip, ip, vp, vp, vp], "hlb200_dev_interp_chroma": [vp, vp, ip, ip, vp, vp, vp, vp],
        "hlb200_dev_tq_recon": [vp, vp, vp, vp, vp, vp, ip, ip, ip, ip, vp, vp, vp, vp, vp],
        "hlb200_dev_interp_luma_batch": [vp, ip, ip, ip, C.c_size_t, vp, vp, vp], "hlb200_dev_interp_chroma_batch": [vp, vp, ip, ip, ip, C.c_size_t, vp, vp, vp, vp],
        "hlb200_dev_tq_recon_batch": [vp, vp, vp, vp, vp, vp, ip, ip, ip, C.c_size_t, ip, ip, vp, vp, vp, vp, vp], "hlb200_dev_sad4x4": [vp, vp, ip, ip, ip, vp, vp],
        "hlb200_dev_svc_inter_recon_batch": [vp, vp, vp, vp, vp, vp, ip, ip, ip, C.c_size_t, ip, ip, vp, vp, vp, vp, vp, vp, vp],
        "hlb200_svc_layer_picture": [vp, ip, ip, ip, ip, vp, vp, vp, vp, vp],
        "hlb200_svc_layer_picture_resampled": [vp, ip, ip, ip, vp, vp, vp, ip, ip, ip, vp],
        "hlb200_svc_layer_picture_resampled_from": [vp, ip, ip, ip, vp, ip, ip, vp],
        "hlb200_dev_svc_derive_motion_batch": [vp, vp, ip, ip, ip, vp, vp, vp, vp],
        "hlb200_svc_layer_picture_derived": [vp, ip, ip, ip, ip, vp, vp, vp, C.POINTER(C.c_int32), vp],
        "hlb200_dev_svc_resample_intra_batch": [vp, vp, vp, ip, ip, vp, vp, vp, ip, ip, ip, C.c_size_t, C.c_size_t, ip, vp],
        "hlb200_dev_svc_bl_recon_batch": [vp, vp, vp, vp, vp, vp, ip, ip, ip, C.c_size_t, ip, ip, vp, vp, vp, vp, vp, vp],
        "hlb200_dev_me_cost": [vp, vp, ip, ip, ip, vp, ip, vp, vp], "hlb200_dev_int_alu_probe": [ip, ip, vp, vp, C.POINTER(C.c_uint64)],
        "hlb200_slice_bits_batch_async": [C.POINTER(vp), C.POINTER(C.c_int32), ip], "hlb200_slice_bits_download": [vp, vp, C.c_size_t, C.POINTER(C.c_uint32)],
        "hlb200_host_register": [vp, C.c_size_t], "hlb200_host_unregister": [vp],
        "hlb200_dev_selftest": [ip, C.c_uint, C.POINTER(ip)], "hlb200_dev_tma_probe": [vp, ip, ip, ip, ip, ip, C.POINTER(ip)],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes, f.restype = args, ip
    lib.hlb200_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        raise Hlb200Error("%s failed with HL_ERROR %d: %s" % (what, rc, load().hlb200_last_error().decode()))


def ptr(a):
    """host pointer of a C-contiguous numpy array"""
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class Stream:
    """One encoder stream context (hlb200_ctx_t)."""

    def __init__(self, width, height, max_refs=1, device=0):
        self.lib = load()
        check(self.lib.hlb200_init(device), "hlb200_init")
        self.w, self.h = width, height
        self.nmb = (width // 16) * (height // 16)
        self.ctx = C.c_void_p()
        check(self.lib.hlb200_stream_create(width, height, max_refs, C.byref(self.ctx)), "hlb200_stream_create")

    def close(self):
        if self.ctx:
            self.lib.hlb200_stream_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _planes(self, yuv):
        w, h = self.w, self.h
        yuv = np.ascontiguousarray(yuv, np.uint8).reshape(-1)
        y, u, v = yuv[:w * h], yuv[w * h:w * h * 5 // 4], yuv[w * h * 5 // 4:w * h * 3 // 2]
        return y, u, v

    def upload_frame(self, yuv):
        y, u, v = self._planes(yuv)
        check(self.lib.hlb200_frame_upload(self.ctx, ptr(y), ptr(u), ptr(v), self.w, self.w // 2), "frame_upload")
        check(self.lib.hlb200_stream_sync(self.ctx), "sync")

    def upload_slot(self, slot, yuv):
        y, u, v = self._planes(yuv)
        check(self.lib.hlb200_slot_upload(self.ctx, slot, ptr(y), ptr(u), ptr(v)), "slot_upload")

    def download_slot(self, slot):
        w, h = self.w, self.h
        out = np.zeros(w * h * 3 // 2, np.uint8)
        y, u, v = out[:w * h], out[w * h:w * h * 5 // 4], out[w * h * 5 // 4:]
        check(self.lib.hlb200_slot_download(self.ctx, slot, ptr(y), ptr(u), ptr(v)), "slot_download")
        return out

    def interp_luma(self, ref_slot, motion):
        out = np.zeros((self.h, self.w), np.uint8)
        check(self.lib.hlb200_interp_luma(self.ctx, ref_slot, ptr(motion), ptr(out)), "interp_luma")
        return out

    def interp_chroma(self, ref_slot, motion):
        u = np.zeros((self.h // 2, self.w // 2), np.uint8)
        v = np.zeros_like(u)
        check(self.lib.hlb200_interp_chroma(self.ctx, ref_slot, ptr(motion), ptr(u), ptr(v)), "interp_chroma")
        return u, v

    def tq_recon(self, qp, pred_yuv, chroma_qp_index_offset=0):
        py, pu, pv = self._planes(pred_yuv)
        coeffs = np.zeros(self.nmb, MB_COEFFS)
        rec = np.zeros(self.w * self.h * 3 // 2, np.uint8)
        w, h = self.w, self.h
        ry, ru, rv = rec[:w * h], rec[w * h:w * h * 5 // 4], rec[w * h * 5 // 4:]
        check(self.lib.hlb200_tq_recon(self.ctx, qp, chroma_qp_index_offset, ptr(py), ptr(pu), ptr(pv), ptr(coeffs), ptr(ry), ptr(ru), ptr(rv)), "tq_recon")
        return coeffs, rec

    def svc_layer_picture(self, qp, motion=None, pred_yuv=None, ref_slot=0, cur_slot=1, chroma_qp_index_offset=0):
        """one picture of an SVC enhancement layer (this context = the layer): base-mode inter macroblocks predicted from ref_slot with `motion`, or -- pred_yuv
        given -- I_BL macroblocks predicted by the host-resampled planes; returns (coefficients, reconstruction = frame store cur_slot)"""
        coeffs = np.zeros(self.nmb, MB_COEFFS)
        if pred_yuv is not None:
            py, pu, pv = self._planes(pred_yuv)
            rc = self.lib.hlb200_svc_layer_picture(self.ctx, -1, cur_slot, qp, chroma_qp_index_offset, None, ptr(py), ptr(pu), ptr(pv), ptr(coeffs))
        else:
            m = np.ascontiguousarray(motion)
            rc = self.lib.hlb200_svc_layer_picture(self.ctx, ref_slot, cur_slot, qp, chroma_qp_index_offset, ptr(m), None, None, None, ptr(coeffs))
        check(rc, "svc_layer_picture")
        return coeffs, self.download_slot(cur_slot)

    def svc_layer_picture_derived(self, qp, base, geom, ref_slot=0, cur_slot=1, chroma_qp_index_offset=0):
        """a P picture of an SVC enhancement layer, inter-layer motion derivation on the device too: base = the reference layer's macroblock fields (SVC_BASE_MB),
        geom = SVC_GEOM[1]; returns (coefficients, reconstruction, derived motion, status bits); a refused picture (status != 0) returns (None, None, motion, status)"""
        coeffs, motion, st = np.zeros(self.nmb, MB_COEFFS), np.zeros(self.nmb, MB_MOTION), C.c_int32(0)
        b, g = np.ascontiguousarray(base), np.ascontiguousarray(geom)
        rc = self.lib.hlb200_svc_layer_picture_derived(self.ctx, ref_slot, cur_slot, qp, chroma_qp_index_offset, ptr(b), ptr(g), ptr(motion), C.byref(st), ptr(coeffs))
        if rc == 7 and st.value:
            return None, None, motion, st.value
        check(rc, "svc_layer_picture_derived")
        return coeffs, self.download_slot(cur_slot), motion, 0

    def svc_layer_picture_resampled(self, qp, ref_layer_yuv, ref_w, ref_h, cur_slot=1, chroma_qp_index_offset=0, level_idc=0):
        """an I picture of an SVC enhancement layer, Intra_Base resampling on the device too: ref_layer_yuv = the reference layer's reconstruction (tight Y|U|V)"""
        coeffs = np.zeros(self.nmb, MB_COEFFS)
        r = np.ascontiguousarray(ref_layer_yuv, np.uint8)
        ys, cs = ref_w * ref_h, ref_w * ref_h // 4
        ry, ru, rv = np.ascontiguousarray(r[:ys]), np.ascontiguousarray(r[ys:ys + cs]), np.ascontiguousarray(r[ys + cs:ys + 2 * cs])
        check(self.lib.hlb200_svc_layer_picture_resampled(self.ctx, cur_slot, qp, chroma_qp_index_offset, ptr(ry), ptr(ru), ptr(rv), ref_w, ref_h, level_idc, ptr(coeffs)),
              "svc_layer_picture_resampled")
        return coeffs, self.download_slot(cur_slot)

    def svc_layer_picture_resampled_from(self, qp, ref_stream, ref_stream_slot, cur_slot=1, chroma_qp_index_offset=0, level_idc=0):
        """the same with the reference layer's reconstruction taken from frame store ref_stream_slot of another Stream (same or another GPU)"""
        coeffs = np.zeros(self.nmb, MB_COEFFS)
        check(self.lib.hlb200_svc_layer_picture_resampled_from(self.ctx, cur_slot, qp, chroma_qp_index_offset, ref_stream.ctx, ref_stream_slot, level_idc, ptr(coeffs)),
              "svc_layer_picture_resampled_from")
        return coeffs, self.download_slot(cur_slot)

    def sad4x4(self, pred_y, satd=False):
        out = np.zeros((self.h // 4, self.w // 4), np.int32)
        p = np.ascontiguousarray(pred_y, np.uint8)
        check(self.lib.hlb200_sad4x4(self.ctx, ptr(p), int(satd), ptr(out)), "sad4x4")
        return out

    def homogeneity8x8(self):
        out = np.zeros((self.h // 8, self.w // 8), np.int32)
        check(self.lib.hlb200_homogeneity8x8(self.ctx, ptr(out)), "homogeneity8x8")
        return out

    def me_cost(self, ref_slot, qp, cands):
        out = np.zeros(len(cands), ME_COST)
        check(self.lib.hlb200_me_cost(self.ctx, ref_slot, qp, ptr(cands), len(cands), ptr(out)), "me_cost")
        return out

    def slice_status(self):
        st = np.zeros(16, np.int32)
        rc = self.lib.hlb200_slice_status(self.ctx, ptr(st))
        if rc != 0:
            raise Hlb200Error("slice kernel watchdog fired: head %d tail %d total %d code %d dbg %s" % (st[0], st[1], st[2], st[3], st[4:12].tolist()))
        return st

    def slice_encode(self, params):
        rec = np.zeros(self.nmb, MB_RECORD)
        check(self.lib.hlb200_slice_encode(self.ctx, C.byref(params), ptr(rec)), "slice_encode")
        self.slice_status()
        return rec


class Encoder:
    """Host-side mirror of the reference's per-stream encode flow for the device path (source/h264/hl_codec_264.c:404-1006):
    frame 0 (and every gop_size-th) is an IDR picture, the rest are P pictures predicted from the previous reconstructions;
    fixed QP (rate control stays on the host).  Frame stores rotate like a sliding-window DPB with `refs` (= max_ref_frame) pictures.

    The reference's encoder writes num_ref_idx_l0_active_minus1 = 0 into every slice header it creates (hl_codec_264_slice.c:289-291)
    and the search loop runs over num_ref_idx_l0_active_minus1 + 1 list entries (hl_codec_264_rdo.c:845,866), so whatever
    max_ref_frame is, a P picture is searched in RefPicList0[0] = the previous reconstruction only (traced: hl_ref_driver --refs 4
    visits refIdx 0 alone and reconstructs exactly like --refs 1).  `active_refs` is that count; raise it only to exercise the
    kernel's multi-reference loop, for which the reference offers no behaviour to compare with."""

    def __init__(self, width, height, qp=31, me_range=16, refs=1, gop_size=400, device=0, active_refs=1, early_term=0, deblock=0):
        self.st = Stream(width, height, refs, device)
        self.early_term, self.deblock = early_term, deblock
        self.qp, self.me_range, self.refs, self.gop = qp, me_range, refs, gop_size
        self.active_refs = active_refs
        self.order = []      # slots holding reference pictures, most recent first
        self.n = 0

    def params(self):
        p = SliceParams()
        idr = (self.n % self.gop) == 0
        if idr:
            self.order = []
        p.slice_type = 0 if idr else 1
        p.qp, p.me_range, p.chroma_qp_index_offset = self.qp, self.me_range, 0
        p.me_early_term_flag, p.deblock_flag = self.early_term, self.deblock
        p.num_refs = min(len(self.order), self.refs, self.active_refs) if not idr else 0
        p.cur_slot = next(s for s in range(self.refs + 1) if s not in self.order)
        for i, s in enumerate(self.order[:p.num_refs]):
            p.ref_slot[i] = s
        return p

    def advance(self, p):
        self.order.insert(0, p.cur_slot)
        del self.order[self.refs:]
        self.n += 1

    def encode(self, yuv, want_recon=False):
        p = self.params()
        self.st.upload_frame(yuv)
        rec = self.st.slice_encode(p)
        recon = self.st.download_slot(p.cur_slot) if want_recon else None
        self.advance(p)
        return rec, recon

    def close(self):
        self.st.close()


def encode_batch(encoders, frames):
    """one picture of every encoder in a single launch (hlb200_slice_encode_batch_async); returns after the launch is queued"""
    lib = load()
    n = len(encoders)
    ps = (SliceParams * n)()
    ctxs = (C.c_void_p * n)()
    for i, (e, f) in enumerate(zip(encoders, frames)):
        p = e.params()
        C.memmove(C.byref(ps[i]), C.byref(p), C.sizeof(SliceParams))
        ctxs[i] = e.st.ctx
        if f is not None:
            e.st.upload_frame(f)
    check(lib.hlb200_slice_encode_batch_async(ctxs, ps, n), "slice_encode_batch_async")
    for i, e in enumerate(encoders):
        e.advance(ps[i])
    return ps
