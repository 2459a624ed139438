"""Host mirrors of the evaluate/ steps that sit directly on the hot path; the Tester harness itself
(COCO loading, JSON writing, plotting) is out of scope."""
from .pipeline import process_batch  # noqa: F401
from .prn_assign import prn_process, prn_process_batch  # noqa: F401
from .tta import crop_with_factor, get_multiplier, get_outputs, handle_heat, multi_scale_flip  # noqa: F401
