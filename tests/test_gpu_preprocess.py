"""GPU parity of spvo_preprocess[_device] (crop + cv::resize INTER_LINEAR + /255, BASE:68-121, NN:139-161) against the
oracle, which tests/test_oracle_preprocess.py pins bit for bit to cv2.resize."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [((375, 1242), (376, 1240)), ((370, 1226), (120, 392)), ((480, 640), (192, 640)), ((376, 1240), (376, 1240)),
         ((256, 512), (128, 256)), ((100, 300), (96, 96)), ((300, 100), (96, 96)), ((31, 17), (200, 120))]


@pytest.mark.parametrize("src,dst", CASES)
def test_preprocess_parity_host_api(spvo, oracle, src, dst):
    rows, cols = src
    H, W = dst
    B = 3
    imgs = np.random.default_rng(rows * 7 + W).integers(0, 256, (B, rows, cols), dtype=np.uint8)
    P = np.arange(B * 12, dtype=np.float32).reshape(B, 3, 4) * np.float32(1.37) + np.float32(100)
    fe = spvo.Frontend(0, 2, 64, 64, 16)
    inp, rs, Pp = fe.preprocess(imgs, H, W, P)
    for b in range(B):
        oi, ors, oP = oracle.preprocess(imgs[b], H, W, P[b])
        assert (rs[b] == ors).all(), (b, int((rs[b] != ors).sum()))
        assert (inp[b].view(np.uint32) == oi.view(np.uint32)).all()
        assert (Pp[b].view(np.uint32) == oP.view(np.uint32)).all()
    fe.close()


def test_preprocess_device_api_with_row_padding(spvo, oracle):
    import torch
    rows, cols, stride, H, W, B = 375, 1242, 1280, 376, 1240, 2
    rng = np.random.default_rng(3)
    padded = rng.integers(0, 256, (B, rows, stride), dtype=np.uint8)
    fe = spvo.Frontend(0, 2, H, W, 16)
    fe.set_stream(torch.cuda.current_stream().cuda_stream)
    d_img = torch.from_numpy(padded).cuda()
    d_in = torch.empty(B, H, W, dtype=torch.float32, device="cuda")
    d_u8 = torch.empty(B, H, W, dtype=torch.uint8, device="cuda")
    P = np.tile(np.array([718.856, 0, 607.1928, -386.1448, 0, 718.856, 185.2157, 0, 0, 0, 1, 0], np.float32), (B, 1))
    P0 = P.copy()
    fe.preprocess_device(d_img, B, rows, cols, stride, H, W, d_in, d_u8, P)
    torch.cuda.synchronize()
    for b in range(B):
        oi, ors, oP = oracle.preprocess(padded[b, :, :cols], H, W, P0[b])
        assert (d_u8[b].cpu().numpy() == ors).all()
        assert (d_in[b].cpu().numpy().view(np.uint32) == oi.view(np.uint32)).all()
        assert (P[b].reshape(3, 4).view(np.uint32) == oP.view(np.uint32)).all()
    # only one of the two outputs requested
    d_in.zero_()
    fe.preprocess_device(d_img, B, rows, cols, stride, H, W, d_in, None, None)
    torch.cuda.synchronize()
    assert (d_in[1].cpu().numpy() == oracle.preprocess(padded[1, :, :cols], H, W)[0]).all()
    fe.close()


def test_preprocess_invalid_arguments_and_mirror(spvo, oracle):
    fe = spvo.Frontend(0, 2, 64, 64, 16)
    with pytest.raises(spvo.SpvoError):
        fe.preprocess(np.zeros((1, 4, 4000), np.uint8), 376, 8)   # the crop would be empty
    with pytest.raises(spvo.SpvoError):
        fe.preprocess(np.zeros((1, 8, 8), np.uint8), 0, 8)
    fe.close()
    # the mirror class hands the image over exactly like NN:139-161
    m = spvo.SuperPointFeatureFrontEnd(model_batch_size=2, input_height=120, input_width=392, max_keypoints=100)
    img = np.random.default_rng(5).integers(0, 256, (370, 1226), dtype=np.uint8)
    P = np.array([[707.09, 0, 601.89, 0], [0, 707.09, 183.11, 0], [0, 0, 1, 0]], np.float32)
    oi, ors, oP = oracle.preprocess(img, 120, 392, P)
    m.preprocessImage(img, P, 1)
    assert (m.input_data_[1] == oi).all() and (m.images_dq[-1] == ors).all() and (P == oP).all()


def test_preprocess_random_geometries(spvo, oracle):
    """40 random (source size, network input size) pairs incl. odd widths (scalar store tail) and heavy up / down
    scaling; one handle, so the coefficient-table cache is exercised across changing geometries."""
    rng = np.random.default_rng(99)
    fe = spvo.Frontend(0, 2, 64, 64, 16)
    done = 0
    for _ in range(60):
        rows, cols = int(rng.integers(9, 400)), int(rng.integers(9, 600))
        H, W = int(rng.integers(8, 300)), int(rng.integers(8, 400))
        try:
            oracle.crop_geometry(rows, cols, H, W)
        except ValueError:
            continue
        imgs = rng.integers(0, 256, (2, rows, cols), dtype=np.uint8)
        inp, rs, _ = fe.preprocess(imgs, H, W)
        for b in range(2):
            oi, ors, _ = oracle.preprocess(imgs[b], H, W)
            assert (rs[b] == ors).all(), ((rows, cols), (H, W), int((rs[b] != ors).sum()))
            assert (inp[b].view(np.uint32) == oi.view(np.uint32)).all()
        done += 1
    assert done >= 40
    fe.close()
