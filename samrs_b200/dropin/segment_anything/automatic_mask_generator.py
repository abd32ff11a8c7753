"""`SamAutomaticMaskGenerator` is exported by the reference package
(`segment_anything/__init__.py:15`) but no SAMRS driver uses it; the name is kept
importable and construction fails loudly."""


class SamAutomaticMaskGenerator:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "SamAutomaticMaskGenerator is outside the box-prompted hot path that samrs_b200 implements")
