"""GPU parity of the picture-batch entry points (hlb200_dev_interp_luma_batch / _chroma_batch / hlb200_dev_tq_recon_batch): one launch over
n pictures must give, picture by picture, what the oracle (interpolation) and the single-picture entry points (transform / quantisation /
reconstruction, themselves pinned against the oracle in test_codec_264_transf.py) give."""
import numpy as np
import pytest

from gpu_util import random_motion
from oracle_lib import load_oracle, oracle_predict_frame

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,n", [(176, 144, 3), (64, 48, 5)])
def test_picture_batch_parity(w, h, n):
    import torch
    from hartallo_b200 import lib as hl
    from test_oracle_pinned import stress_plane
    lib = hl.load()
    rng = np.random.default_rng(w + n)
    ysz, csz = w * h, w * h // 4
    fb = ysz + 2 * csz
    nmb = (w // 16) * (h // 16)

    def frame():
        return np.concatenate([stress_plane(rng, h, w).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1), stress_plane(rng, h // 2, w // 2).reshape(-1)])
    refs = np.stack([frame() for _ in range(n)])
    srcs = np.stack([frame() for _ in range(n)])
    motion = np.concatenate([random_motion(rng, nmb, far_every=4 if i == 1 else 0) for i in range(n)])
    dev = torch.device("cuda:0")
    d_ref, d_src = torch.from_numpy(refs).to(dev), torch.from_numpy(srcs).to(dev)
    d_motion = torch.from_numpy(motion.view(np.uint8)).to(dev)
    d_pred, d_rec = torch.zeros_like(d_ref), torch.zeros_like(d_ref)
    d_coef = torch.zeros(n * nmb * hl.MB_COEFFS.itemsize, dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    b = d_ref.data_ptr()
    p = d_pred.data_ptr()
    hl.check(lib.hlb200_dev_interp_luma_batch(b, w, h, n, fb, d_motion.data_ptr(), p, sp), "interp_luma_batch")
    hl.check(lib.hlb200_dev_interp_chroma_batch(b + ysz, b + ysz + csz, w, h, n, fb, d_motion.data_ptr(), p + ysz, p + ysz + csz, sp), "interp_chroma_batch")
    s, r = d_src.data_ptr(), d_rec.data_ptr()
    qp = 28
    hl.check(lib.hlb200_dev_tq_recon_batch(s, s + ysz, s + ysz + csz, p, p + ysz, p + ysz + csz, w, h, n, fb, qp, 0, d_coef.data_ptr(), r, r + ysz, r + ysz + csz, sp),
             "tq_recon_batch")
    torch.cuda.synchronize()
    pred, rec = d_pred.cpu().numpy(), d_rec.cpu().numpy()
    coef = d_coef.cpu().numpy().view(hl.MB_COEFFS).reshape(n, nmb)
    o = load_oracle()
    st = hl.Stream(w, h, 1)
    for i in range(n):
        oy, ou, ov = oracle_predict_frame(o, refs[i], w, h, motion[i * nmb:(i + 1) * nmb])
        assert np.array_equal(pred[i, :ysz].reshape(h, w), oy), i
        assert np.array_equal(pred[i, ysz:ysz + csz].reshape(h // 2, w // 2), ou) and np.array_equal(pred[i, ysz + csz:].reshape(h // 2, w // 2), ov), i
        st.upload_frame(srcs[i])
        c1, r1 = st.tq_recon(qp, pred[i])
        assert np.array_equal(r1, rec[i]), i
        assert c1.tobytes() == coef[i].tobytes(), i
    st.close()
