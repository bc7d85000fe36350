import numpy as np


def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    err = np.mean((np.asarray(image_true, dtype=np.float64) - np.asarray(image_test, dtype=np.float64)) ** 2)
    if data_range is None:
        data_range = 255.0
    return 10.0 * np.log10((data_range ** 2) / err)


def structural_similarity(*a, **k):
    raise NotImplementedError("skimage stub: SSIM is not used by the e2e harness")
