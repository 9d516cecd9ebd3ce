"""Host mirror of the edge-image functions of the reference's ``src/utils/img_utils.py``, backed by the CUDA library: same names
and argument meaning, so ``exp_mgr.py:343-350`` can call these instead of OpenCV / SciPy.  SURVEY.md 8f rank 3.

    edges = jnp.stack([normalize_to_unit_range(smoothen_edges(image_to_edge(jnp_to_ocv_n255(image)))) for image in images])

becomes ``edges = edge_maps(np.stack([jnp_to_ocv_n255(image) for image in images]), th1, th2)`` - one call, all frames of the window
through the same launches, result left on the device as the ``edges`` operand of ``loss_func`` / ``Plan.set_window``.

The image pre-processing before Canny (``preprocess_image``: non-local-means denoise, CLAHE, sharpen, bilateral filter) is not part
of this library.  There is no CPU fallback: without the built library the import of ``eincm_b200.plan`` fails.
"""
import ctypes as C

import numpy as np

from . import plan as _plan

__all__ = ['jnp_to_ocv_n255', 'image_to_edge', 'smoothen_edges', 'eincm_inv_exp_dist_transform', 'normalize_to_unit_range', 'edge_maps',
           'fast_nl_means_denoising', 'preprocess_image']

_IEDT_ALPHA = 6.0 / 5.541            # configs/edge_extraction/smoothen/iedt.yaml


def jnp_to_ocv_n255(img) -> np.ndarray:
    """src/utils/img_utils.py:44-45: ``(img * 255).astype(uint8)`` (host side; the frames arrive as host arrays from the loaders)."""
    return (np.asarray(img) * 255).astype(np.uint8)


def _u8_frames(images):
    a = np.ascontiguousarray(np.asarray(images))
    if a.dtype != np.uint8:
        raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must be uint8 (what cv.Canny receives); use jnp_to_ocv_n255')
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3 or a.shape[0] < 1 or a.shape[1] < 1 or a.shape[2] < 1:
        raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must have shape (H, W) or (R, H, W)')
    return a


def edge_maps(images, th1=30, th2=80, smoothen='gaussian', k_size=1, alpha=_IEDT_ALPHA, return_canny=False, device=None, stream=None,
              stages=0):
    """Edge images of a window on the device: ``images`` uint8 ``(R, H, W)`` (or ``(H, W)``) -> float64 CUDA tensor of the same shape,
    each image ``normalize_to_unit_range(smoothen(Canny(image)))`` as exp_mgr.py:343-350 stages it.  ``smoothen``: 'gaussian'
    (``smoothen_edges``, gaussian.yaml) or 'iedt' (``eincm_inv_exp_dist_transform``, iedt.yaml).  Asynchronous on the current
    stream; with ``return_canny`` also returns cv.Canny's uint8 image.  ``stages``: mask of ``plan.EINCM_EDGE_STAGE_*`` (0 = all)."""
    import torch
    if isinstance(images, torch.Tensor) and images.is_cuda:
        if images.dtype != torch.uint8:
            raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must be uint8')
        single = images.dim() == 2
        d_img = images.contiguous()
        if single:
            d_img = d_img[None]
        if d_img.dim() != 3:
            raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must have shape (H, W) or (R, H, W)')
    else:
        single = np.asarray(images).ndim == 2
        a = _u8_frames(images)
        dev = f'cuda:{torch.cuda.current_device() if device is None else device}'
        d_img = torch.from_numpy(a).to(dev)
    if smoothen not in ('gaussian', 'iedt'):
        raise _plan.EincmError(_plan.EINCM_EUNSUPPORTED, f"smoothen must be 'gaussian' or 'iedt', not {smoothen!r}")
    n, H, W = (int(v) for v in d_img.shape)
    lib = _plan.load_library()
    p = _plan.EdgeParams(float(th1), float(th2), _plan.EINCM_SMOOTHEN_GAUSSIAN if smoothen == 'gaussian' else _plan.EINCM_SMOOTHEN_IEDT,
                         int(stages), float(k_size), float(alpha))
    wsb = int(lib.eincm_edge_workspace_bytes(H, W, n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=d_img.device)
    out = torch.empty((n, H, W), dtype=torch.float64, device=d_img.device)
    canny = torch.empty((n, H, W), dtype=torch.uint8, device=d_img.device) if return_canny else None
    with torch.cuda.device(d_img.device):
        s = stream if stream is not None else torch.cuda.current_stream()
        rc = lib.eincm_edge_maps(d_img.device.index, d_img.data_ptr(), n, H, W, C.byref(p), out.data_ptr(),
                                 canny.data_ptr() if canny is not None else None, ws.data_ptr(), wsb, int(s.cuda_stream))
        ws.record_stream(s)
        d_img.record_stream(s)
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_edge_maps failed (bad shape, thresholds, sigma or alpha)')
    if single:
        out = out[0]
        canny = canny[0] if canny is not None else None
    return (out, canny) if return_canny else out


def fast_nl_means_denoising(images, h=4, template_win_size=3, search_win_size=11, device=None, stream=None):
    """``cv.fastNlMeansDenoising(img, None, h, template_win_size, search_win_size)`` (the first step of ``preprocess_image``,
    src/utils/img_utils.py:147-157) for uint8 frames ``(H, W)`` or ``(R, H, W)``, bit-exact with OpenCV: uint8 CUDA tensor of the same
    shape.  Asynchronous on the current stream."""
    import torch
    if isinstance(images, torch.Tensor) and images.is_cuda:
        if images.dtype != torch.uint8:
            raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must be uint8')
        single = images.dim() == 2
        d_img = images.contiguous()[None] if single else images.contiguous()
    else:
        single = np.asarray(images).ndim == 2
        dev = f'cuda:{torch.cuda.current_device() if device is None else device}'
        d_img = torch.from_numpy(_u8_frames(images)).to(dev)
    n, H, W = (int(v) for v in d_img.shape)
    lib = _plan.load_library()
    wsb = int(lib.eincm_nlm_workspace_bytes(int(template_win_size), int(search_win_size)))
    if wsb == 0:
        raise _plan.EincmError(_plan.EINCM_EUNSUPPORTED, 'template window <= 15 and search window <= 41 are implemented')
    ws = torch.empty(wsb, dtype=torch.uint8, device=d_img.device)
    out = torch.empty_like(d_img)
    with torch.cuda.device(d_img.device):
        s = stream if stream is not None else torch.cuda.current_stream()
        rc = lib.eincm_nlm_denoise(d_img.device.index, d_img.data_ptr(), n, H, W, float(h), int(template_win_size), int(search_win_size),
                                   out.data_ptr(), ws.data_ptr(), wsb, int(s.cuda_stream))
        ws.record_stream(s)
        d_img.record_stream(s)
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_nlm_denoise failed')
    return out[0] if single else out


def _device_frames(images, device=None):
    """uint8 frames ``(H, W)`` or ``(R, H, W)`` on the host or the device -> (contiguous CUDA tensor ``(n, H, W)``, was a single frame)."""
    import torch
    if isinstance(images, torch.Tensor) and images.is_cuda:
        if images.dtype != torch.uint8:
            raise _plan.EincmError(_plan.EINCM_EINVAL, 'frames must be uint8')
        single = images.dim() == 2
        return (images.contiguous()[None] if single else images.contiguous()), single
    single = np.asarray(images).ndim == 2
    dev = f'cuda:{torch.cuda.current_device() if device is None else device}'
    return torch.from_numpy(_u8_frames(images)).to(dev), single


def clahe_apply(images, clip_limit=5, tile_grid_size=(10, 10), device=None, stream=None):
    """``cv.createCLAHE(clipLimit=clip_limit, tileGridSize=tile_grid_size).apply(img)`` (src/utils/img_utils.py:159-161) for uint8 frames
    ``(H, W)`` or ``(R, H, W)``, bit-exact with OpenCV: uint8 CUDA tensor of the same shape.  Asynchronous on the current stream."""
    import torch
    d_img, single = _device_frames(images, device)
    n, H, W = (int(v) for v in d_img.shape)
    tx, ty = int(tile_grid_size[0]), int(tile_grid_size[1])
    lib = _plan.load_library()
    wsb = int(lib.eincm_clahe_workspace_bytes(n, tx, ty))
    if wsb == 0:
        raise _plan.EincmError(_plan.EINCM_EINVAL, 'the tile grid must be at least 1 x 1')
    ws = torch.empty(wsb, dtype=torch.uint8, device=d_img.device)
    out = torch.empty_like(d_img)
    with torch.cuda.device(d_img.device):
        s = stream if stream is not None else torch.cuda.current_stream()
        rc = lib.eincm_clahe(d_img.device.index, d_img.data_ptr(), n, H, W, float(clip_limit), tx, ty, out.data_ptr(), ws.data_ptr(), wsb,
                             int(s.cuda_stream))
        ws.record_stream(s)
        d_img.record_stream(s)
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_clahe failed')
    return out[0] if single else out


def sharpen(images, sigma=3, alpha=1.5, beta=-0.5, gamma=0.0, return_blur=False, device=None, stream=None):
    """The sharpening step of ``preprocess_image`` (src/utils/img_utils.py:163-178): ``blur = cv.GaussianBlur(img, None, sigma, ...)`` on
    the uint8 frame, then ``cv.addWeighted(img, alpha, blur, beta, gamma)``; bit-exact with OpenCV.  uint8 CUDA tensor(s)."""
    import torch
    d_img, single = _device_frames(images, device)
    n, H, W = (int(v) for v in d_img.shape)
    lib = _plan.load_library()
    out = torch.empty_like(d_img)
    blur = torch.empty_like(d_img) if return_blur else None
    with torch.cuda.device(d_img.device):
        s = stream if stream is not None else torch.cuda.current_stream()
        rc = lib.eincm_sharpen(d_img.device.index, d_img.data_ptr(), n, H, W, float(sigma), float(alpha), float(beta), float(gamma),
                               blur.data_ptr() if return_blur else None, out.data_ptr(), int(s.cuda_stream))
        d_img.record_stream(s)
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_sharpen failed (Gaussian kernels of up to 63 taps: sigma <= 10)')
    if return_blur:
        return (out[0], blur[0]) if single else (out, blur)
    return out[0] if single else out


def preprocess_image(img, denoise_h=4, denoise_template_win_size=3, denoise_search_win_size=11, clahe_clip_limit=5,
                     clahe_tile_grid_size=(10, 10), sharpen_kernel_size=3, sharpen_sigma_x=2, sharpen_alpha=1.5, sharpen_beta=-0.5,
                     bilateral_filter_neigh_diameter=5, bilateral_filter_sigma_color=15, bilateral_filter_sigma_space=15) -> np.ndarray:
    """src/utils/img_utils.py:131-191 with the non-local-means denoise (99 % of its run time with OpenCV: ~110 ms of ~112 ms per
    640x480 frame on 8 host threads), CLAHE and the Gaussian sharpen on the device, chained there (one copy in, one out); the bilateral
    filter stays an OpenCV call, made exactly as the reference makes it (it needs ``cv2``, like the reference does: in the opencv-python
    build its output is a property of the build, see include/eincm.h).  uint8 in (or a [0, 1] float image like the reference accepts),
    uint8 out, identical to the reference's result.  ``sharpen_sigma_x`` is accepted and unused, as in the reference's call: OpenCV reads
    ``cv.GaussianBlur(img, None, kernel_size, sigma_x, sigma_y)`` as sigmaX = kernel_size, dst = sigma_x."""
    import cv2 as cv
    del sharpen_sigma_x
    a = np.asarray(img)
    if a.dtype != np.uint8:
        a = jnp_to_ocv_n255(a)
    d_img = fast_nl_means_denoising(a, denoise_h, denoise_template_win_size, denoise_search_win_size)
    clahe_img = clahe_apply(d_img, clahe_clip_limit, clahe_tile_grid_size)
    sharp = sharpen(clahe_img, sharpen_kernel_size, sharpen_alpha, sharpen_beta).cpu().numpy()
    return cv.bilateralFilter(sharp, bilateral_filter_neigh_diameter, bilateral_filter_sigma_color, bilateral_filter_sigma_space)


def image_to_edge(img, apert_size=3, th1=30, th2=80) -> np.ndarray:
    """src/utils/img_utils.py:194-211 (``cv.Canny(img, th1, th2, None, 3, True)``): uint8 0 / 255 image, bit-exact."""
    if int(apert_size) != 3:
        raise _plan.EincmError(_plan.EINCM_EUNSUPPORTED, 'only the 3x3 Sobel aperture of the shipped configs is implemented')
    _, canny = edge_maps(img, th1, th2, return_canny=True)
    return canny.cpu().numpy()


def smoothen_edges(edge_img, k_size=1, sigma=1) -> np.ndarray:
    """src/utils/img_utils.py:213-222 on a uint8 image (what ``image_to_edge`` returns).  As OpenCV executes the reference's call,
    ``k_size`` is the Gaussian sigma and ``sigma`` lands on the ``dst`` parameter (ignored)."""
    del sigma
    out = edge_maps(edge_img, smoothen='gaussian', k_size=k_size, stages=_plan.EINCM_EDGE_STAGE_SMOOTHEN)
    return out.cpu().numpy()


def eincm_inv_exp_dist_transform(edge_img, alpha=6) -> np.ndarray:
    """src/utils/img_utils.py:231-235 (non-zero pixels are edges)."""
    a = (np.asarray(edge_img) != 0).astype(np.uint8) * 255
    out = edge_maps(a, smoothen='iedt', alpha=alpha, stages=_plan.EINCM_EDGE_STAGE_SMOOTHEN)
    return out.cpu().numpy()


def normalize_to_unit_range(a):
    """src/utils/img_utils.py:24-25 on a CUDA tensor or array (float64)."""
    import torch
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).cuda()
    t = t.to(torch.float64)
    mn, mx = t.min(), t.max()
    return (t - mn) / (mx - mn + float(np.finfo(np.float64).eps))
