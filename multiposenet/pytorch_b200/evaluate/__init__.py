"""Host mirrors of the evaluate/ steps that sit directly on the hot path; the Tester harness itself
(COCO loading, JSON writing, plotting) is out of scope."""
from .pipeline import process_batch  # noqa: F401
from .prn_assign import prn_process, prn_process_batch  # noqa: F401
