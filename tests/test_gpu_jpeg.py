"""GPU tier: frame ingest (svb_jpeg_decode_host, svb_scan_batch_v1_jpeg_host) through the C ABI: the decoded frames against
cv2.imdecode's (the committed golden, tests/golden/jpeg.npz, and fresh encodes where cv2 is importable), and the compressed
whole-path call against the device path run on the decoded frames."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_jpeg_decode_golden(scanner, golden):
    from svb200 import Scanner

    g = golden("jpeg")
    names = [k[:-5] for k in g if k.endswith("_file") and not k.startswith(("unsupported", "malformed"))]
    for nm in names:
        h, w, _ = (int(v) for v in g[nm + "_shape"])
        blob, offs = Scanner.pack_jpegs([g[nm + "_file"].tobytes()] * 3)  # a batch of three copies
        out, status = scanner.jpeg_decode(blob, offs, h, w, want_status=True)
        out = out.cpu().numpy()
        assert int(status.sum()) == 0, nm
        for i in range(3):
            assert hashlib.sha256(np.ascontiguousarray(out[i]).tobytes()).digest() == g[nm + "_sha"].tobytes(), (nm, i)
    blob, offs = Scanner.pack_jpegs([g["unsupported_progressive_file"].tobytes()])
    with pytest.raises(NotImplementedError):
        scanner.jpeg_decode(blob, offs, 270, 480)
    blob, offs = Scanner.pack_jpegs([g["malformed_truncated_file"].tobytes()])
    with pytest.raises(ValueError):
        scanner.jpeg_decode(blob, offs, 270, 480)
    blob, offs = Scanner.pack_jpegs([g["frame270_q90_r4_444_file"].tobytes()])
    with pytest.raises(ValueError):  # geometry mismatch
        scanner.jpeg_decode(blob, offs, 540, 960)


def test_jpeg_decode_vs_cv2_1080p_batch(scanner):
    cv2 = pytest.importorskip("cv2")
    from svb200 import Scanner
    from svb200 import frames as F

    files, want = [], []
    for i in range(6):
        im = F.add_noise_host(F.make_frame(8800 + i, 1080, 1920).image, 40 + i)
        params = [cv2.IMWRITE_JPEG_QUALITY, (80, 90, 95)[i % 3]]
        if i != 4:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, (8, 120, 1, 30, 0, 15)[i]]
        ok, buf = cv2.imencode(".jpg", im, params)
        files.append(buf.tobytes())
        want.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
    blob, offs = Scanner.pack_jpegs(files)
    out, status = scanner.jpeg_decode(blob, offs, 1080, 1920, want_status=True)
    out = out.cpu().numpy()
    assert int(status.sum()) == 0
    for i in range(6):
        assert np.array_equal(out[i], want[i]), i


def test_scan_batch_jpeg_host_equals_device_path_on_decoded_frames(scanner, golden):
    """compressed frames in, boards out == the device path on the frames the decoder produced (multi-chunk not needed: the
    chunking is the raw host path's)"""
    import torch
    from svb200 import Scanner

    g = golden("jpeg")
    files = [g["frame540_q90_r8_420_file"].tobytes(), g["frame540_q95_r60_420_file"].tobytes(), g["frame540_q75_r0_420_file"].tobytes()] * 7
    blob, offs = Scanner.pack_jpegs(files)
    frames = scanner.jpeg_decode(blob, offs, 540, 960)
    dev = scanner.scan_batch(frames)
    torch.cuda.synchronize()
    host = scanner.scan_batch_jpeg_host(blob, offs, 540, 960)
    for k in ("digits", "corners", "found", "conf"):
        assert np.array_equal(host[k], dev[k].cpu().numpy()), k
    assert int(host["found"].sum()) == len(files)
