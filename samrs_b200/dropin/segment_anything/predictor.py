"""`SamPredictor` over the samrs_b200 engine (reference: segment_anything/predictor.py:17-271).

Same stateful surface: `set_image` caches the image embedding on the device, `predict_torch` decodes a
batch of prompts against it.  Inputs and outputs are torch tensors on the model's CUDA device, exactly as
the drivers use them (`main_sam_hbox_semantic.py:155,174-189`)."""
from typing import Optional, Tuple

import numpy as np
import torch

from .modeling import Sam
from .utils.transforms import ResizeLongestSide


class SamPredictor:
    def __init__(self, sam_model: Sam) -> None:
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    def set_image(self, image: np.ndarray, image_format: str = "RGB") -> None:
        assert image_format in ["RGB", "BGR"], f"image_format must be in ['RGB', 'BGR'], is {image_format}."
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        # resize to long side 1024: identity for 1024-long tiles; otherwise the original pixels go to the device and
        # `samrs_resize_bilinear_u8` reproduces PIL's resampler bit for bit (the reference does it on the host with PIL)
        target = self.transform.get_preprocess_shape(image.shape[0], image.shape[1], self.transform.target_length)
        # H2D: asynchronous when the caller's array lives in pinned host memory (the copy is then ordered on the current
        # stream like every other call of this class), the reference's blocking copy otherwise
        host = torch.from_numpy(np.ascontiguousarray(image))
        input_image_torch = host.to(self.device, non_blocking=host.is_pinned())
        if tuple(target) != tuple(image.shape[:2]):
            if input_image_torch.dtype != torch.uint8 or input_image_torch.dim() != 3 or input_image_torch.shape[2] != 3:
                raise NotImplementedError("samrs_b200 resizes 8-bit HWC RGB images; got " + str(tuple(image.shape)) + " " + str(image.dtype))
            input_image_torch = self.model._require_engine().resize_image(input_image_torch, target)
        # the engine reads HWC uint8 directly; the permute of the reference is a layout detail of its conv
        self._set_device_image(input_image_torch, hwc=True, original_image_size=image.shape[:2])

    @torch.no_grad()
    def set_torch_image(self, transformed_image: torch.Tensor, original_image_size: Tuple[int, ...]) -> None:
        assert (
            len(transformed_image.shape) == 4
            and transformed_image.shape[1] == 3
            and max(*transformed_image.shape[2:]) == self.model.image_encoder.img_size
        ), f"set_torch_image input must be BCHW with long side {self.model.image_encoder.img_size}."
        assert transformed_image.shape[0] == 1, "one image at a time"
        self._set_device_image(transformed_image[0], hwc=False, original_image_size=original_image_size)

    def _set_device_image(self, img: torch.Tensor, hwc: bool, original_image_size) -> None:
        engine = self.model._require_engine()
        self.reset_image()
        self.original_size = tuple(original_image_size)
        self.input_size = tuple(img.shape[:2]) if hwc else tuple(img.shape[-2:])
        if img.dtype != torch.uint8:
            rounded = img.round().clamp(0, 255)
            if not torch.equal(rounded, img.to(rounded.dtype)):
                raise NotImplementedError("samrs_b200 encodes 8-bit images; got non-integral pixel values")
            img = rounded.to(torch.uint8)
        self.features = engine.encode(img.to(self.device), chw=not hwc)
        self._engine_features, self._engine_gen = self.features, engine.feature_gen
        self.is_image_set = True

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None,
                multimask_output: bool = True, return_logits: bool = False):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        coords_torch, labels_torch, box_torch, mask_input_torch = None, None, None, None
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied if point_coords is supplied."
            point_coords = self.transform.apply_coords(point_coords, self.original_size)
            coords_torch = torch.as_tensor(point_coords, dtype=torch.float, device=self.device)[None, :, :]
            labels_torch = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)[None, :]
        if box is not None:
            box = self.transform.apply_boxes(box, self.original_size)
            box_torch = torch.as_tensor(box, dtype=torch.float, device=self.device)[None, :]
        if mask_input is not None:
            mask_input_torch = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)[None, :, :, :]
        masks, iou_predictions, low_res_masks = self.predict_torch(
            coords_torch, labels_torch, box_torch, mask_input_torch, multimask_output, return_logits=return_logits)
        return (masks[0].detach().cpu().numpy(), iou_predictions[0].detach().cpu().numpy(),
                low_res_masks[0].detach().cpu().numpy())

    @torch.no_grad()
    def predict_torch(self, point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = True, return_logits: bool = False,
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        engine = self.model._require_engine()
        # the engine caches ONE image: re-install ours if the embedding was assigned by hand or if
        # another predictor on the same Sam has encoded its own image since (the reference keeps features per predictor)
        if self.features is not self._engine_features or engine.feature_gen != self._engine_gen:
            engine.set_features(self.features)
            self._engine_features, self._engine_gen = self.features, engine.feature_gen
        if boxes is not None and boxes.dim() == 1:
            boxes = boxes[None, :]
        low_res_masks, iou_predictions = engine.decode(
            boxes=boxes, point_coords=point_coords, point_labels=point_labels, mask_input=mask_input,
            multimask_output=multimask_output)
        masks = engine.postprocess(low_res_masks, self.input_size, self.original_size, return_logits=return_logits)
        return masks, iou_predictions, low_res_masks

    def get_image_embedding(self) -> torch.Tensor:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        assert self.features is not None, "Features must exist if an image has been set."
        return self.features

    @property
    def device(self) -> torch.device:
        return self.model.device

    def reset_image(self) -> None:
        self.is_image_set = False
        self.features = None
        self._engine_features = None
        self._engine_gen = -1
        self.orig_h = None
        self.orig_w = None
        self.input_h = None
        self.input_w = None
