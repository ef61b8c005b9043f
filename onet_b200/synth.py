"""On-device synthesis of the training frames — host-side mirror of the reference generator
`Rayleigh_bg_Gaussian_EOT_generator_20230208.py`:

    get_rayleigh_frame(snr)            :219-249  ->  get_rayleigh_frames(n, snr, ...)   (n frames per call, on the device)
    get_k_frame(snr)                   :178-217  ->  get_k_frames(n, snr, ...)          (compound-Gaussian K clutter, uncorrelated)
    add_gaussian_template_on_clutter_v3 :62-176  ->  add_gaussian_targets(frames, cx, cy, w, h, theta, snr)
    prepare_frames / prepare_data      :251-321  ->  prepare_data(...): the reference's dataset dictionary
                                                     {'<type>_imgs', '<type>_labels', 'psnr', 'desc'} (torch.save-able, read by
                                                     dataloader/simbg4onet_20230209.py:298-305)

The clutter background and the compositing of the targets are CUDA kernels (csrc/synth.cuh).  The per-target scalars
(window position and size, quadratic form of the rotated Gaussian) are derived on the host in float64 from
(cx, cy, w, h, theta) with the reference's formulas; like the reference, a target whose window leaves the frame or an snr
outside its table raises ValueError.  The target parameters are drawn with numpy exactly as the reference draws them
(normal(centre, 30 / 24), normal(10, 2), normal(18, 2), rand * 180).  There is no CPU path.
"""
import numpy as np
import torch

from ._lib import call, ptr

SNR_LIST = [12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, -1, -2]          # reference :114
_TARGET_DTYPE = np.dtype([("lx", "<i4"), ("ly", "<i4"), ("wr", "<i4"), ("hr", "<i4"),
                          ("a", "<f4"), ("b", "<f4"), ("c", "<f4"), ("thr", "<f4")])


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _need_cuda(dev):
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise RuntimeError("onet_b200.synth has no CPU path")
    return dev


def rayleigh_background(n, h, w, seed=1981, sigma=1.0, device="cuda", stream_id=0):
    """[n,h,w] fp32 Rayleigh(sigma) amplitudes (reference :221, rayleigh.rvs(loc=0, scale=1))."""
    dev = _need_cuda(device)
    out = torch.empty(n, h, w, dtype=torch.float32, device=dev)
    call("onet_synth_rayleigh", ptr(out), out.numel(), float(sigma), int(seed), int(stream_id), _stream(dev))
    return out


def k_background(n, h, w, seed=1981, nu=5, device="cuda", stream_id=0):
    """[n,h,w] fp32 K-distributed amplitudes, shape parameter nu (reference get_k_frame uses gamma_shape=5, :189):
    Rayleigh speckle x sqrt(Gamma(nu, 1/nu)) texture, spatially uncorrelated (see csrc/synth.cuh)."""
    dev = _need_cuda(device)
    out = torch.empty(n, h, w, dtype=torch.float32, device=dev)
    call("onet_synth_kclutter", ptr(out), out.numel(), int(nu), int(seed), int(stream_id), _stream(dev))
    return out


def target_table(cx, cy, w, h, theta, img_h, img_w, host_threshold=False):
    """Per-target records for onet_synth_add_targets from arrays of shape [frames, targets] (float64, reference formulas)."""
    cx, cy, w, h, theta = (np.asarray(v, dtype=np.float64) for v in (cx, cy, w, h, theta))
    sigma_x = (w / 2 - 0.5) / 2                                          # :68-69
    sigma_y = (h / 2 - 0.5) / 2
    wr = np.int32(sigma_x * 2.5 + 0.5)                                   # gaussian_kernel2d :36-37
    hr = np.int32(sigma_y * 2.5 + 0.5)
    th = -1 * theta                                                      # :45 (the reference feeds degrees to cos / sin as is)
    a = np.cos(th) ** 2 / (2 * sigma_x ** 2) + np.sin(th) ** 2 / (2 * sigma_y ** 2)
    b = -np.sin(2 * th) / (4 * sigma_x ** 2) + np.sin(2 * th) / (4 * sigma_y ** 2)
    c = np.sin(th) ** 2 / (2 * sigma_x ** 2) + np.cos(th) ** 2 / (2 * sigma_y ** 2)
    h_t, w_t = 2 * hr + 1, 2 * wr + 1
    ly = (cy - (h_t - 1) / 2).astype(np.int64)                           # int() truncates toward zero, :76-79
    ry = (cy + (h_t - 1) / 2).astype(np.int64)
    lx = (cx - (w_t - 1) / 2).astype(np.int64)
    rx = (cx + (w_t - 1) / 2).astype(np.int64)
    if np.any((ly < 0) | (lx < 0) | (ry > img_h) | (rx > img_w)) or np.any((ly + h_t > img_h) | (lx + w_t > img_w)):
        raise ValueError('template location is beyond the image boundaries!')        # :82-83
    tab = np.zeros(cx.shape, dtype=_TARGET_DTYPE)
    thr = np.full(cx.shape, -1.0)          # negative: the kernel reduces kgauss.std() over the window itself (:142)
    if host_threshold:                     # the same number on the host in float64 (tests)
        for idx in np.ndindex(cx.shape):
            KX, KY = np.meshgrid(np.arange(-wr[idx], wr[idx] + 1), np.arange(-hr[idx], hr[idx] + 1))
            kg = np.exp(-(a[idx] * KX ** 2 + 2 * b[idx] * KX * KY + c[idx] * KY ** 2))
            thr[idx] = kg.max() - 2 * kg.std()
    tab["lx"], tab["ly"], tab["wr"], tab["hr"] = lx, ly, wr, hr
    tab["a"], tab["b"], tab["c"], tab["thr"] = a, b, c, thr
    return tab


def add_gaussian_targets(frames, cx, cy, w, h, theta, snr):
    """`add_gaussian_template_on_clutter_v3` (swerling type 0) for every target of every frame, in order.
    frames: [n,H,W] fp32 CUDA tensor, modified in place; cx..theta: [n, targets] arrays.  Returns (frames, mask [n,H,W] bool,
    erc [n] = mean(bg^2) of the untouched backgrounds)."""
    if not frames.is_cuda:
        raise RuntimeError("onet_b200.synth has no CPU path")
    SNR_LIST.index(snr)                                                  # ValueError like the reference's snr_lis.index(snr)
    assert frames.dim() == 3 and frames.dtype == torch.float32 and frames.is_contiguous()
    n, H, W = frames.shape
    tab = target_table(cx, cy, w, h, theta, H, W)
    assert tab.shape[0] == n
    T = tab.shape[1]
    dtab = torch.from_numpy(tab.view(np.uint8).reshape(n, T * _TARGET_DTYPE.itemsize).copy()).to(frames.device)
    masks = torch.zeros(n, H, W, dtype=torch.uint8, device=frames.device)
    erc = torch.empty(n, dtype=torch.float32, device=frames.device)
    call("onet_synth_add_targets", ptr(frames), ptr(masks), n, H, W, ptr(dtab), T, float(snr), ptr(erc), _stream(frames.device))
    return frames, masks.bool(), erc


def draw_targets(n, img_sz, target_num=20, rng=None):
    """The reference's draws for one frame (:232-237), for n frames: centres around the frame centre, w ~ N(10,2), h ~ N(18,2)."""
    rng = np.random if rng is None else rng
    cx0, cy0 = img_sz[0] / 2, img_sz[1] / 2
    cx = rng.normal(cx0, 30, (n, target_num))
    cy = rng.normal(cy0, 24, (n, target_num))
    w = rng.normal(10, 2, (n, target_num))
    h = rng.normal(18, 2, (n, target_num))
    theta = rng.rand(n, target_num) * 180
    return cx, cy, w, h, theta


def _frames(kind, n, snr, img_sz, target_num, seed, device, rng):
    bg = (rayleigh_background if kind == "rayleigh" else k_background)(n, img_sz[0], img_sz[1], seed=seed, device=device)
    rng = np.random.RandomState(seed) if rng is None else rng
    return add_gaussian_targets(bg, *draw_targets(n, img_sz, target_num, rng), snr)[:2]


def get_rayleigh_frames(n, snr=10, img_sz=(400, 400), target_num=20, seed=1981, device="cuda", rng=None):
    """n frames of `get_rayleigh_frame(snr)`: (frames [n,H,W] fp32, fg masks [n,H,W] bool), on the device."""
    return _frames("rayleigh", n, snr, img_sz, target_num, seed, device, rng)


def get_k_frames(n, snr=10, img_sz=(400, 400), target_num=20, seed=1981, device="cuda", rng=None):
    """n frames of `get_k_frame(snr)` with the uncorrelated K-clutter background."""
    return _frames("kdist", n, snr, img_sz, target_num, seed, device, rng)


def prepare_data(img_sz=(224, 224), bg_type="rayleigh", file_name=None, fnums=150, snrs=range(0, 11), seed=1981, device="cuda"):
    """The reference's dataset dictionary (prepare_data :288-321): per snr `fnums` frames of 400 x 400, every frame scaled to
    [0,1] by its own min / max (array_normal, :263), centre-cropped to img_sz, labels as float32.  Saved with torch.save when
    file_name is given — the file dataloader/simbg4onet_20230209.py:298-305 loads."""
    from .evaluate import normalize_per_frame
    imgs, labels, psnrs = [], [], []
    for k, snr in enumerate(snrs):
        get = get_rayleigh_frames if bg_type == "rayleigh" else get_k_frames
        f, m = get(fnums, snr=snr, seed=seed + 7919 * k, device=device)
        f = normalize_per_frame(f.unsqueeze(1))
        top, left = (f.shape[2] - img_sz[0]) // 2, (f.shape[3] - img_sz[1]) // 2
        imgs.append(f[:, :, top:top + img_sz[0], left:left + img_sz[1]].contiguous())
        labels.append(m[:, top:top + img_sz[0], left:left + img_sz[1]].float())
        psnrs.extend([snr] * fnums)
    data = {f"{bg_type}_imgs": torch.cat(imgs).cpu(), f"{bg_type}_labels": torch.cat(labels).cpu(), "psnr": psnrs,
            "desc": f"{bg_type} clutter add 20 extended targets [pure fg higher than mu-2*simga] in each frame; "
                    f"{fnums} frames per snr, synthesised on the GPU by onet_b200.synth"}
    if file_name is not None:
        torch.save(data, file_name)
    return data


def save_checkpoint(onet, epoch, path):
    """The reference's checkpoint file (Train_Onet_on_simclutter_20250407.py:264-266): {'net': state_dict, 'epoch': epoch}."""
    torch.save({"net": {k: v.detach().cpu() for k, v in onet.state_dict().items()}, "epoch": epoch}, path)


def load_checkpoint(onet, path):
    """Loads a checkpoint written by the reference training script (or by save_checkpoint) — :493, :523."""
    ck = torch.load(path, map_location=lambda storage, loc: storage)
    onet.load_state_dict(ck["net"])
    return ck.get("epoch")
