"""On-device synthesis of the training frames — host-side mirror of the reference generator
`Rayleigh_bg_Gaussian_EOT_generator_20230208.py`:

    get_rayleigh_frame(snr)            :219-249  ->  get_rayleigh_frames(n, snr, ...)   (n frames per call, on the device)
    get_k_frame(snr)                   :178-217  ->  get_k_frames(n, snr, ...)          (correlated K field, or uncorrelated)
    generate_K_distributed_noise (K_distributed_SeaClutter_Simulation_20210919.py:469-526) -> k_correlated_background
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
    call("onet_synth_rayleigh", ptr(out), out.numel(), float(sigma), int(seed), int(stream_id), _stream(dev), device=dev)
    return out


def k_background(n, h, w, seed=1981, nu=5, device="cuda", stream_id=0):
    """[n,h,w] fp32 K-distributed amplitudes, shape parameter nu (reference get_k_frame uses gamma_shape=5, :189):
    Rayleigh speckle x sqrt(Gamma(nu, 1/nu)) texture, spatially uncorrelated (see csrc/synth.cuh)."""
    dev = _need_cuda(device)
    out = torch.empty(n, h, w, dtype=torch.float32, device=dev)
    call("onet_synth_kclutter", ptr(out), out.numel(), int(nu), int(seed), int(stream_id), _stream(dev), device=dev)
    return out


_KFIELD_CONST = {}


def _kfield_constants(size, v, dev):
    """Per (size, v): the Gamma-field autocorrelation of eq. (69) of Tough & Ward as the reference samples it (:477-484) and the
    square root of the speckle power spectrum |f|^-0.6 (:276-295), float64 on the device."""
    key = (size, v, str(dev))
    if key not in _KFIELD_CONST:
        xs = np.linspace(10, size, num=size, endpoint=True)
        XS, YS = np.meshgrid(xs, xs)
        acf = 1 + np.exp(-(XS + YS) / 10) * np.cos(np.pi * YS / 8) / v
        f = np.linspace(0.1, size / 10.0, num=size, endpoint=True)
        Fx, Fy = np.meshgrid(f, f)
        spec = np.sqrt(np.sqrt(Fx ** 2 + Fy ** 2) ** (-0.6))
        _KFIELD_CONST[key] = (torch.from_numpy(acf).to(dev), torch.from_numpy(spec).to(dev))
    return _KFIELD_CONST[key]


def normal_white(n, h, w, seed=1981, device="cuda", stream_id=0):
    """[n,h,w] float64 standard normal white noise (Philox + Box-Muller on the device)."""
    dev = _need_cuda(device)
    out = torch.empty(n, h, w, dtype=torch.float64, device=dev)
    call("onet_synth_normal", ptr(out), out.numel(), int(seed), int(stream_id), _stream(dev), device=dev)
    return out


def mnlt(x, v):
    """`mnlt(x, v)` of the reference (:83-91) on the device: Gamma(shape v, scale 1) quantile of Phi(x); float64, integer v."""
    if not x.is_cuda:
        raise RuntimeError("onet_b200.synth has no CPU path")
    x = x.contiguous().double()
    y = torch.empty_like(x)
    call("onet_kfield_mnlt", ptr(x), x.numel(), int(v), ptr(y), _stream(x.device), device=x.device)
    return y


def k_correlated_background(n, size, v=5, seed=1981, device="cuda", white=None, return_texture=False):
    """n square fields of the reference's correlated K-distributed clutter, `generate_K_distributed_noise(size, size, v)`
    (K_distributed_SeaClutter_Simulation_20210919.py:469-526), on the device: correlated Gamma texture (white noise -> measured
    polynomial coefficients -> per-element quadratic root -> coloured Gaussian field -> mnlt) times a correlated complex speckle
    field, amplitude as fp32 [n,size,size].  The element-wise stages are this library's kernels; the two FFT pairs go through
    torch.fft (cuFFT) in complex128.  `white` = (texture noise, speckle noise), float64 [n,size,size], replays given draws
    (tests use the reference's own); otherwise both are generated on the device from `seed`."""
    dev = _need_cuda(device)
    if white is None:
        w1 = normal_white(n, size, size, seed=seed, device=dev, stream_id=0)
        w2 = normal_white(n, size, size, seed=seed, device=dev, stream_id=1)
    else:
        w1, w2 = (t.to(dev).double().contiguous() for t in white)
        assert w1.shape == (n, size, size) and w2.shape == (n, size, size)
    st = _stream(dev)
    acf, spec = _kfield_constants(size, v, dev)
    per = size * size
    g = mnlt(w1, v)
    sums = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    call("onet_kfield_coeff_sums", ptr(w1), ptr(g), n, per, ptr(sums), st, device=dev)
    # alpha_n = S_n^2 / (pi n! 2^n), normalised by alpha_0 (:131-137, :488): [a, b] = [alpha_2, alpha_1] / alpha_0
    alpha = sums ** 2 / (np.pi * torch.tensor([1.0, 2.0, 8.0], dtype=torch.float64, device=dev))
    coeffs = torch.stack([alpha[:, 2] / alpha[:, 0], alpha[:, 1] / alpha[:, 0]], dim=1).contiguous()
    roots = torch.empty(n, size, size, dtype=torch.complex128, device=dev)
    call("onet_kfield_acf_root", ptr(coeffs), ptr(acf), n, per, ptr(roots), st, device=dev)
    gcn = torch.fft.ifft2(torch.fft.fft2(w1) * torch.sqrt(torch.fft.fft2(roots))).real.contiguous()     # :495-497
    texture = mnlt(gcn, v)
    speckle = torch.fft.ifft2(torch.fft.fft2(w2) * spec).contiguous()                                    # :287-296
    amp = torch.empty(n, size, size, dtype=torch.float32, device=dev)
    call("onet_kfield_amplitude", ptr(speckle), ptr(texture), n * per, ptr(amp), st, device=dev)
    return (amp, texture) if return_texture else amp


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
    call("onet_synth_add_targets", ptr(frames), ptr(masks), n, H, W, ptr(dtab), T, float(snr), ptr(erc), _stream(frames.device), device=frames.device)
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
    if kind == "rayleigh":
        bg = rayleigh_background(n, img_sz[0], img_sz[1], seed=seed, device=device)
    elif kind == "kdist_correlated":
        assert img_sz[0] == img_sz[1], "the reference's K field is square (its speckle field is size x size, :514)"
        bg = k_correlated_background(n, img_sz[0], v=5, seed=seed, device=device)
    else:
        bg = k_background(n, img_sz[0], img_sz[1], seed=seed, device=device)
    rng = np.random.RandomState(seed) if rng is None else rng
    return add_gaussian_targets(bg, *draw_targets(n, img_sz, target_num, rng), snr)[:2]


def get_rayleigh_frames(n, snr=10, img_sz=(400, 400), target_num=20, seed=1981, device="cuda", rng=None):
    """n frames of `get_rayleigh_frame(snr)`: (frames [n,H,W] fp32, fg masks [n,H,W] bool), on the device."""
    return _frames("rayleigh", n, snr, img_sz, target_num, seed, device, rng)


def get_k_frames(n, snr=10, img_sz=(400, 400), target_num=20, seed=1981, device="cuda", rng=None, correlated=True):
    """n frames of `get_k_frame(snr)` (:178-217): the reference's correlated K field (gamma_shape 5) + 20 extended targets;
    correlated=False uses the spatially uncorrelated compound-Gaussian background instead (one kernel, no FFT)."""
    return _frames("kdist_correlated" if correlated else "kdist", n, snr, img_sz, target_num, seed, device, rng)


def assemble_dataset(per_snr, img_sz=(224, 224), bg_type="rayleigh"):
    """The reference's dataset dictionary (prepare_data :288-321) from per-snr groups [(frames [n,1,H,W] in [0,1], masks [n,H,W],
    snr), ...]: centre crop to img_sz (transforms.CenterCrop, :296), images float32 [N,1,h,w], labels float32 [N,h,w], one psnr
    entry per frame.  Pure tensor bookkeeping on whatever device the groups live on; the result is on the CPU."""
    imgs, labels, psnrs = [], [], []
    for f, m, snr in per_snr:
        top, left = int(round((f.shape[2] - img_sz[0]) / 2.0)), int(round((f.shape[3] - img_sz[1]) / 2.0))
        imgs.append(f[:, :, top:top + img_sz[0], left:left + img_sz[1]].contiguous().float().cpu())
        labels.append(m[:, top:top + img_sz[0], left:left + img_sz[1]].float().cpu())
        psnrs.extend([snr] * f.shape[0])
    n_per = per_snr[0][0].shape[0] if per_snr else 0
    return {f"{bg_type}_imgs": torch.cat(imgs), f"{bg_type}_labels": torch.cat(labels), "psnr": psnrs,
            "desc": f"{bg_type} clutter add 20 extended targets [pure fg higher than mu-2*simga] in each frame; "
                    f"{n_per} frames per snr, synthesised on the GPU by onet_b200.synth"}


def prepare_data(img_sz=(224, 224), bg_type="rayleigh", file_name=None, fnums=150, snrs=range(0, 11), seed=1981, device="cuda"):
    """The reference's dataset dictionary (prepare_data :288-321): per snr `fnums` frames of 400 x 400, every frame scaled to
    [0,1] by its own min / max (array_normal, :263), centre-cropped to img_sz, labels as float32.  Saved with torch.save when
    file_name is given — the file dataloader/simbg4onet_20230209.py:298-305 loads."""
    from .evaluate import normalize_per_frame
    groups = []
    for k, snr in enumerate(snrs):
        get = get_rayleigh_frames if bg_type == "rayleigh" else get_k_frames
        f, m = get(fnums, snr=snr, seed=seed + 7919 * k, device=device)
        groups.append((normalize_per_frame(f.unsqueeze(1)), m, snr))
    data = assemble_dataset(groups, img_sz, bg_type)
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
