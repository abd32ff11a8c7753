"""samrs_b200: Blackwell-native SAM box-prompted mask engine behind SAMRS's segment_anything surface."""
import os

DROPIN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
__all__ = ["DROPIN_PATH"]
