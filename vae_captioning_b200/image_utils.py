"""Image file -> uint8 RGB [224, 224, 3] (reference: utils/image_utils.py:5-13): cv2 decode, resize (bilinear, cv2's
default), BGR->RGB, grey images replicated to three channels. The array is what the VGG16 entry points take
(`vc_vgg_forward_u8`, `vc_train_step_images_u8`); mean subtraction happens on the device."""
import numpy as np


def load_image(image_path, shape=(224, 224)):
    import cv2
    img = cv2.imread(image_path)
    if img is None:
        raise FileNotFoundError(image_path)
    img = cv2.resize(img, shape)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    if img.ndim == 2:
        img = np.repeat(img[:, :, None], 3, axis=2)
    return img
