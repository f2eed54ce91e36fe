"""CPU: the C-ABI library loads and exports every symbol include/nervecl.h declares, the ctypes binding
agrees with the header's arity, and the host-side mirror keeps the reference's interface contract.
No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_decls():
    src = open(os.path.join(ROOT, "include", "nervecl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct.*?\} nervecl_conv_params;", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(nervecl_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_library_exports_every_declared_symbol():
    from nerve_cl_b200 import _lib
    lib = _lib.load()
    decls = header_decls()
    assert len(decls) >= 40
    for name, nargs in decls.items():
        assert hasattr(lib, name), f"{name} declared in nervecl.h but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib._SIGNATURES[name]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib._SIGNATURES[name])}"
    assert set(_lib._SIGNATURES) == set(decls)
    assert lib.nervecl_abi_version() == 7
    assert lib.nervecl_error_string(-2).decode().startswith("misaligned")


def test_conv_params_struct_layout():
    from nerve_cl_b200._lib import ConvParams
    # 16 x 4-byte scalars then 8-byte aligned pointer/pitch pairs (matches the C struct on LP64)
    assert ConvParams.x.offset == 64
    assert ctypes.sizeof(ConvParams) == 64 + 8 * 12 + 24 + 8 + 16  # ABI v2: + x2, ldx2, Cin2, x2_center; v3: + colsum; v7: + sign_bits, sign_mode
    assert ConvParams.x2.offset == 160 and ConvParams.Cin2.offset == 176 and ConvParams.colsum.offset == 184


def test_torch_library_ops_registered():
    from nerve_cl_b200 import ops
    for name in ops.OP_NAMES:
        assert hasattr(torch.ops.nervecl, name)
    assert {"conv2d_fwd", "conv2d_wgrad", "warp_fwd", "warp_bwd", "corr_fwd", "corr_bwd", "dwconv3x3_fwd",
            "tfuse_fwd", "ewc_penalty_fwd", "ewc_fisher_accum"} <= set(ops.OP_NAMES)


def test_no_cpu_fallback():
    from nerve_cl_b200.models import SuperResolutionNet, warp_features
    from nerve_cl_b200.continual import EWC
    m = SuperResolutionNet(num_features=16, num_residual_blocks=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.rand(1, 3, 3, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        warp_features(torch.rand(1, 8, 4, 4), torch.zeros(1, 2, 4, 4))
    lin = torch.nn.Linear(4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        EWC(lin).register_task(0, [(torch.rand(2, 4), torch.rand(2, 4))])
    with pytest.raises((RuntimeError, NotImplementedError)):
        torch.ops.nervecl.fill_zero(torch.zeros(4))


REFERENCE_KEYS_TINY = [
    "feature_extractor.head.0.weight", "feature_extractor.head.0.bias",
    "feature_extractor.body.0.depthwise.weight", "feature_extractor.body.0.pointwise.weight",
    "feature_extractor.body.0.bn.weight", "feature_extractor.body.0.bn.bias",
    "feature_extractor.body.0.bn.running_mean", "feature_extractor.body.0.bn.running_var",
    "feature_extractor.body.0.bn.num_batches_tracked",
    "motion_estimator.flow_net.0.weight", "motion_estimator.flow_net.6.bias",
    "temporal_aggregator.attention.0.weight", "temporal_aggregator.attention.4.bias",
    "temporal_aggregator.refine.channel_attention.fc.0.weight",
    "temporal_aggregator.refine.channel_attention.fc.2.weight",
    "temporal_aggregator.refine.spatial_attention.conv.weight",
    "residual_blocks.0.layers.0.0.weight", "residual_blocks.0.layers.4.0.bias", "residual_blocks.0.lff.weight",
    "gff.0.weight", "gff.0.bias", "upsampler.conv.weight", "upsampler.conv.bias",
]


def test_module_contract_matches_reference():
    """Constructor, attributes, parameter count and state_dict keys/shapes of reference super_resolution.py:279-325."""
    from nerve_cl_b200.models import SuperResolutionNet
    m = SuperResolutionNet()
    assert (m.scale_factor, m.temporal_window, m.num_frames) == (2, 1, 3)
    assert m.get_num_parameters() == 1_987_283                     # SURVEY.md section 6 (measured on the reference)
    assert m.conv_macs_per_pixel() == 2_201_634                    # SURVEY.md section 8a closed form
    sd = m.state_dict()
    assert len([k for k in sd]) == 131 + 9                         # 131 parameters + 3 x (mean, var, counter)
    for k in REFERENCE_KEYS_TINY:
        assert k in sd, k
    assert tuple(sd["motion_estimator.flow_net.0.weight"].shape) == (128, 81, 3, 3)
    assert tuple(sd["temporal_aggregator.attention.0.weight"].shape) == (64, 192, 3, 3)
    assert tuple(sd["residual_blocks.7.lff.weight"].shape) == (64, 224, 1, 1)
    assert tuple(sd["upsampler.conv.weight"].shape) == (12, 64, 3, 3)
    m4 = SuperResolutionNet(scale_factor=4, temporal_window=2)
    assert m4.num_frames == 5 and m4.get_num_parameters() == 2_082_937
    assert m.get_flops((128, 128)) > 0
    # checkpoints round-trip by key
    m2 = SuperResolutionNet()
    m2.load_state_dict(sd)


def test_ewc_api_surface():
    from nerve_cl_b200.continual import EWC, OnlineEWC
    lin = torch.nn.Linear(3, 3)
    e = EWC(lin, ewc_lambda=10.0, mode="separate", decay=0.5)
    assert e.penalty() == 0.0 and isinstance(e.penalty(), float)   # python float before any task (ewc.py:210,232)
    assert set(e.state_dict()) == {"ewc_lambda", "mode", "decay", "num_tasks", "fisher_dict", "optpar_dict",
                                   "task_fisher", "task_optpar"}
    o = OnlineEWC(lin)
    assert o.mode == "online" and o.ewc_lambda == 5000.0 and o.decay == 0.999 and o.num_tasks == 0
