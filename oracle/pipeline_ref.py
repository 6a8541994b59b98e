"""numpy restatement of the reference's tile pipeline -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, line by line:
  apply_tta / transform_prob            DigiPathAI/helpers/utils.py:487-501, 505-522
  TissueMaskGenerationOS                DigiPathAI/helpers/utils.py:336-354
  BinMorphoProcessMaskOS                DigiPathAI/helpers/utils.py:200-219
  WSIStridedPatchDataset                DigiPathAI/loaders/dataloader.py:238-312 (grid), 340-390 (tile)
  get_prediction                        DigiPathAI/Segmentation.py:65-189
  getSegmentation (numeric part)        DigiPathAI/Segmentation.py:288-300, 310-322, 336-337, 356

The slide is any object with OpenSlide's ``level_dimensions / level_downsamples / read_region`` (OpenSlide is
not installed here); ``read_region`` must return an RGB array or PIL image of shape [h, w, 3].
skimage's ``rgb2hsv`` / ``threshold_otsu`` are restated (skimage is not installed); cv2 IS installed and is
called exactly as the reference calls it.  np.memmap planes become in-memory arrays of the same dtype.
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- utils.py:487-522
def apply_tta(imgs, tta):
    for i, img in enumerate(imgs):
        if tta == 'FLIP_LEFT_RIGHT':
            img = np.fliplr(img)
        elif tta == 'ROTATE_90':
            img = np.rot90(img)
        elif tta == 'ROTATE_180':
            img = np.rot90(img, 2)
        elif tta == 'ROTATE_270':
            img = np.rot90(img, 3)
        imgs[i] = img
    return imgs


def transform_prob(imgs, tta):
    for i, img in enumerate(imgs):
        if tta == 'FLIP_LEFT_RIGHT':
            img = np.fliplr(img)
        elif tta == 'ROTATE_90':
            img = np.rot90(img, 3)
        elif tta == 'ROTATE_180':
            img = np.rot90(img, 2)
        elif tta == 'ROTATE_270':
            img = np.rot90(img, 1)
        imgs[i] = img
    return imgs


# ----------------------------------------------------------------------------- skimage restatements
def _otsu(image, nbins=256):
    a = np.asarray(image)
    if np.issubdtype(a.dtype, np.integer):
        lo, hi = int(a.min()), int(a.max())
        hist = np.bincount(a.ravel().astype(np.int64) - lo, minlength=hi - lo + 1).astype(float)
        centers = np.arange(lo, hi + 1).astype(float)
    else:
        hist, edges = np.histogram(a.ravel(), nbins)
        hist = hist.astype(float)
        centers = (edges[1:] + edges[:-1]) / 2.
    weight1 = np.cumsum(hist)
    weight2 = np.cumsum(hist[::-1])[::-1]
    with np.errstate(all='ignore'):
        mean1 = np.cumsum(hist * centers) / weight1
        mean2 = (np.cumsum((hist * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[:-1][np.argmax(variance12)]


def _rgb2hsv_s(rgb_u8):
    arr = rgb_u8 / 255.0
    out_v = arr.max(-1)
    delta = np.ptp(arr, -1)
    with np.errstate(all='ignore'):
        out_s = delta / out_v
    out_s[delta == 0.] = 0.
    out_s[np.isnan(out_s)] = 0
    return out_s


def _rgb(region):
    return np.asarray(region.convert('RGB') if hasattr(region, 'convert') else region)


def TissueMaskGenerationOS(slide_obj, level, RGB_min=50):
    img_RGB = slide_obj.read_region((0, 0), level, slide_obj.level_dimensions[level])
    img_RGB = np.transpose(_rgb(img_RGB), axes=[1, 0, 2])
    background_R = img_RGB[:, :, 0] > _otsu(img_RGB[:, :, 0])
    background_G = img_RGB[:, :, 1] > _otsu(img_RGB[:, :, 1])
    background_B = img_RGB[:, :, 2] > _otsu(img_RGB[:, :, 2])
    tissue_RGB = np.logical_not(background_R & background_G & background_B)
    S = _rgb2hsv_s(img_RGB)
    tissue_S = S > _otsu(S)
    min_R = img_RGB[:, :, 0] > RGB_min
    min_G = img_RGB[:, :, 1] > RGB_min
    min_B = img_RGB[:, :, 2] > RGB_min
    return tissue_S & tissue_RGB & min_R & min_G & min_B


def BinMorphoProcessMaskOS(mask, level):
    import cv2
    close_kernel = np.ones((20, 20), dtype=np.uint8)
    image_close = cv2.morphologyEx(np.array(mask), cv2.MORPH_CLOSE, close_kernel)
    open_kernel = np.ones((5, 5), dtype=np.uint8)
    image_open = cv2.morphologyEx(np.array(image_close), cv2.MORPH_OPEN, open_kernel)
    if level <= 2:
        kernel = np.ones((60, 60), dtype=np.uint8)
    elif level == 3:
        kernel = np.ones((35, 35), dtype=np.uint8)
    elif level == 4:
        kernel = np.ones((10, 10), dtype=np.uint8)
    else:
        raise ValueError("Kernel for this level not fixed")
    return cv2.dilate(image_open, kernel, iterations=1)


# ----------------------------------------------------------------------------- dataloader.py:191-390
class WSIStridedPatchDataset:
    def __init__(self, slide, image_size=256, normalize=True, sampling_stride=16, roi_masking=True, mask=None):
        self._slide = slide
        self._image_size = image_size
        self._normalize = normalize
        self._sampling_stride = sampling_stride
        self._roi_masking = roi_masking
        self._given_mask = mask
        self._preprocess()

    def _preprocess(self):
        self._level = len(self._slide.level_dimensions) - 1
        self._sampling_stride = self._sampling_stride // int(self._slide.level_downsamples[self._level])
        X_slide, Y_slide = self._slide.level_dimensions[0]
        factor = self._sampling_stride
        if self._given_mask is not None:
            self._mask = self._given_mask
        else:
            self._mask = TissueMaskGenerationOS(self._slide, self._level)
        self._mask = BinMorphoProcessMaskOS(np.uint8(self._mask), self._level)
        X_mask, Y_mask = self._mask.shape
        if X_slide // X_mask != Y_slide // Y_mask:
            raise Exception('Slide/Mask dimension does not match')
        self._resolution = np.round(X_slide * 1.0 / X_mask)
        if not np.log2(self._resolution).is_integer():
            raise Exception('Resolution (X_slide / X_mask) is not power of 2 : {}'.format(self._resolution))
        self._strided_mask = np.ones_like(self._mask)
        ones_mask = np.zeros_like(self._mask)
        ones_mask[::factor, ::factor] = self._strided_mask[::factor, ::factor]
        if self._roi_masking:
            self._strided_mask = ones_mask * self._mask
        else:
            self._strided_mask = ones_mask
        self._X_idcs, self._Y_idcs = np.where(self._strided_mask)
        self._idcs_num = len(self._X_idcs)

    def __len__(self):
        return self._idcs_num

    def __getitem__(self, idx):
        x_coord, y_coord = self._X_idcs[idx], self._Y_idcs[idx]
        x_max_dim, y_max_dim = self._slide.level_dimensions[0]
        x = int(x_coord * self._resolution - self._image_size // 2)
        y = int(y_coord * self._resolution - self._image_size // 2)
        x = max(0, min(x, x_max_dim - self._image_size))
        y = max(0, min(y, y_max_dim - self._image_size))
        img = np.transpose(_rgb(self._slide.read_region((x, y), 0, (self._image_size, self._image_size))), [1, 0, 2])
        img = np.array(img, dtype=np.float32)
        if self._normalize:
            img = (img - 128.0) / 128.0
        return img, x, y


def _batches(dataset, batch_size):
    """torch DataLoader(batch_size, drop_last=True), sequential sampler (Segmentation.py:92)."""
    n = len(dataset) // batch_size
    for b in range(n):
        items = [dataset[b * batch_size + i] for i in range(batch_size)]
        yield (np.stack([it[0] for it in items]), np.array([it[1] for it in items]),
               np.array([it[2] for it in items]))


# ----------------------------------------------------------------------------- Segmentation.py:65-189
def get_prediction(slide, batch_size=64, models=None, tta_list=None, patch_size=256, stride_size=256,
                   status=None, mask=None, normalise=True):
    dataset_obj = WSIStridedPatchDataset(slide, image_size=patch_size, normalize=True,
                                         sampling_stride=stride_size, roi_masking=True, mask=mask)
    n_batches = len(dataset_obj) // batch_size
    if tta_list is None:
        tta_list = np.array(['DEFAULT'])
    else:
        tta_list = np.concatenate([np.array(['DEFAULT']), np.array(tta_list)])
    dims = slide.level_dimensions[0]
    probs_map = {'mean': np.zeros(dims, dtype=np.float32), 'var': np.zeros(dims, dtype=np.float32)}
    count_map = np.zeros(dims, dtype=np.uint8)
    for ii, (image_patches, x_coords, y_coords) in enumerate(_batches(dataset_obj, batch_size)):
        if status is not None:
            status['progress'] = int(ii * 100.0 / (len(models.keys()) * n_batches))
        patch_predictions = []
        for tta_ in tta_list:
            image_patches = apply_tta(image_patches, tta_)
            for model_name in models.keys():
                prediction = models[model_name].predict(image_patches, batch_size=batch_size, verbose=0, steps=None)
                prediction_trans = transform_prob(prediction, tta_)
                patch_predictions.append(prediction_trans)
        patch_predictions = np.array(patch_predictions)
        m = np.mean(patch_predictions, axis=0)      # hoisted: the reference recomputes these inside the i-loop
        v = np.var(patch_predictions, axis=0)       # (Segmentation.py:167,170); same values every time
        for i in range(batch_size):
            shape = patch_predictions[0, 0].shape
            xs, ys = x_coords[i], y_coords[i]
            probs_map['mean'][xs: xs + shape[0], ys: ys + shape[1]] += m[i, :, :, 1]
            probs_map['var'][xs: xs + shape[0], ys: ys + shape[1]] += v[i, :, :, 1]
            count_map[xs: xs + shape[0], ys: ys + shape[1]] += np.ones_like(patch_predictions[0, 0, :, :, 1], dtype=np.uint8)
    if normalise:
        np.place(count_map, count_map == 0, 1)
        probs_map['mean'] /= count_map
        probs_map['var'] /= count_map ** 2.0
    probs_map['count'] = count_map
    probs_map['dataset'] = dataset_obj
    return slide, probs_map


def getSegmentation(slide, models, patch_size=256, stride_size=128, batch_size=32, tta_list=None, status=None,
                    mask=None):
    """Numeric part of getSegmentation: returns (thresholded mean {0,255} float32 [W,H], probs mean, var)."""
    threshold = 0.3
    _, probs_map = get_prediction(slide, batch_size=batch_size, tta_list=tta_list, models=models,
                                  patch_size=patch_size, stride_size=stride_size, status=status, mask=mask)
    probs = probs_map['mean'].copy()
    np.place(probs_map['mean'], probs_map['mean'] >= threshold, 255)
    np.place(probs_map['mean'], probs_map['mean'] < threshold, 0)
    return np.array(probs_map['mean']), probs, probs_map['var']
