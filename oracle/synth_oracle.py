"""CPU oracle of the frame synthesis — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-numpy (float64) restatement of the reference generator
/root/reference/source_code/Rayleigh_bg_Gaussian_EOT_generator_20230208.py:
    gaussian_kernel2d                      :28-60   (bnorm=False as the caller passes)
    add_gaussian_template_on_clutter_v3    :62-176  (swerling_type 0, the only type get_*_frame uses, :204,241)
    get_rayleigh_frame's target loop       :219-249 (given the background and the target parameters)
Pinned against the unmodified reference functions by tests/golden/make_synth_golden.py -> tests/golden/synth.npz
(tests/test_oracle_golden.py::test_synth_oracle_matches_reference_golden).  Only tests/ may import this."""
import numpy as np

SNR_LIST = [12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, -1, -2]        # :114, snr_lis.index(snr) raises for anything else


def gaussian_kernel2d(sigma_x, sigma_y, theta):                         # :28-60, bnorm=False
    wr = np.int32(sigma_x * 2.5 + 0.5)
    hr = np.int32(sigma_y * 2.5 + 0.5)
    KX, KY = np.meshgrid(np.arange(-wr, wr + 1), np.arange(-hr, hr + 1))
    theta = -1 * theta
    a = np.cos(theta) ** 2 / (2 * sigma_x ** 2) + np.sin(theta) ** 2 / (2 * sigma_y ** 2)
    b = -np.sin(2 * theta) / (4 * sigma_x ** 2) + np.sin(2 * theta) / (4 * sigma_y ** 2)
    c = np.sin(theta) ** 2 / (2 * sigma_x ** 2) + np.cos(theta) ** 2 / (2 * sigma_y ** 2)
    return np.exp(-(a * KX ** 2 + 2 * b * KX * KY + c * KY ** 2))


def add_target(cx, cy, w, h, theta, erc, snr, bg, mask):                # :62-176, swerling 0
    """Composites one target into bg (float64 [H,W], modified in place) and returns (bg, mask > 0)."""
    sigma_x = (w / 2 - 0.5) / 2
    sigma_y = (h / 2 - 0.5) / 2
    kg = gaussian_kernel2d(sigma_x, sigma_y, theta)
    h_t, w_t = kg.shape
    ly, ry = int(cy - (h_t - 1) / 2), int(cy + (h_t - 1) / 2)
    lx, rx = int(cx - (w_t - 1) / 2), int(cx + (w_t - 1) / 2)
    img_h, img_w = bg.shape
    if ly < 0 or lx < 0 or ry > img_h or rx > img_w:
        raise ValueError('template location is beyond the image boundaries!')
    SNR_LIST.index(snr)                                                  # ValueError for an snr outside the table (:122)
    roi = bg[ly:ly + h_t, lx:lx + w_t]
    template = kg * np.sqrt(np.power(10, (snr / 10)) * erc)             # kcoef_peak :90, swerling 0 :96-98
    mask_t = kg > (kg.max() - 2 * kg.std())                             # :142
    bg[ly:ly + h_t, lx:lx + w_t] = (template > roi) * template + roi    # :143-145
    mask[ly:ly + h_t, lx:lx + w_t] += mask_t                            # :153
    return bg, mask > 0


def composite_frame(bg, cx, cy, w, h, theta, snr):                      # get_rayleigh_frame :221-249 after the draws
    bg = np.array(bg, dtype=np.float64)
    erc = np.sum(bg ** 2) / bg.size
    mask = np.zeros_like(bg)
    for i in range(len(cx)):
        bg, mask = add_target(cx[i], cy[i], w[i], h[i], theta[i], erc, snr, bg, mask)
    return bg, mask, erc
