"""Seeded synthetic weight sets and the parameter inventory (TEST INFRASTRUCTURE view).

The generator itself is plain data code shared with bench.py and lives in multiposenet/pytorch_b200/synthetic.py;
tests/test_oracle.py pins `param_spec` against the live reference's state_dict.
"""
from multiposenet.pytorch_b200.synthetic import (BLOCKS, CLS_OUT_GAIN, REG_OUT_GAIN, fpn_spec, make_weights,  # noqa: F401
                                                 param_spec, to_torch_state_dict)
