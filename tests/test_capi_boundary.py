"""The C-ABI library loads and exports every symbol include/boxpath.h declares; argument structs have the header's
layout; without a GPU the library refuses to create a handle (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from tf_eager_object_detection_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'boxpath.h')


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bx_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_documented_entry_points():
    names = header_functions()
    for must in ('bx_create', 'bx_destroy', 'bx_last_error', 'bx_version', 'bx_decode_clip', 'bx_nms', 'bx_proposals',
                 'bx_crop_and_resize', 'bx_roi_pool', 'bx_fpn_assign_levels', 'bx_fpn_roi_features', 'bx_pairwise_iou',
                 'bx_anchor_target', 'bx_proposal_target', 'bx_post_ops_prediction', 'bx_c4_proposal_roi',
                 'bx_c4_proposal_roi_host'):
        assert must in names


def test_library_exports_every_header_symbol_and_binding_covers_them():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('libboxpath.so not built (run __graft_entry__.build())')
    lib = _lib.load()
    names = header_functions()
    for n in names:
        assert hasattr(lib, n), 'libboxpath.so does not export %s' % n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes binding declares argtypes for all of them
    assert lib.bx_version() == 100


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for path in ('model/region_proposal.py:37-81', 'utils/bbox_transform.py:32-55', 'utils/bbox_tf.py:59-78',
                 'model/roi_pooling.py', 'model/fpn/base_fpn_model.py:303-324', 'model/fpn/base_fpn_model.py:152-161',
                 'utils/bbox_tf.py:37-56', 'model/anchor_target.py:29-107', 'model/proposal_target.py:32-124',
                 'model/prediction.py:103-163'):
        assert path in src, path


def test_param_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.ProposalParams) == 4 * 4 + 4 * 4 + 6 * 4
    assert ctypes.sizeof(_lib.AnchorTargetParams) == 2 * 4 + 2 * 4 + 8 * 4 + 2 * 4
    assert ctypes.sizeof(_lib.ProposalTargetParams) == 5 * 4 + 8 * 4
    assert ctypes.sizeof(_lib.PredictionParams) == 8 * 4 + 8 * 4


def test_no_cpu_fallback():
    """Without a CUDA device bx_create fails loudly and the Python layer raises; with one this test is vacuous."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('libboxpath.so not built')
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.BoxpathError) as e:
        _lib.handle(0)
    assert 'no CPU fallback' in str(e.value)
    import numpy as np
    from tf_eager_object_detection_b200 import ops
    with pytest.raises(Exception):
        ops.pairwise_iou(np.zeros((2, 4), np.float32), np.zeros((2, 4), np.float32))


def test_nvtx_ranges_are_optional_and_harmless_without_a_profiler():
    """BX_NVTX=1 brackets every entry point with an NVTX range (header-only nvtx3: resolves its injection library
    lazily, a no-op when no tool is attached).  No GPU needed: an argument error path runs through the guard."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip('libboxpath.so not built')
    import subprocess, sys
    code = ('from tf_eager_object_detection_b200 import _lib\n'
            'lib = _lib.load()\n'
            'for _ in range(3):\n'
            '    assert lib.bx_stats(None, None, 0) == -1\n'
            '    assert lib.bx_reserve(None, 0, 0, 0, None) != 0\n'
            'print(lib.bx_last_error().decode())\n')
    env = dict(os.environ, BX_NVTX='1', PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert 'NULL' in out.stdout


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'tf_eager_object_detection_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'boxpath_ref' not in txt, f
