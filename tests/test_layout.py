"""CPU: the drop-in modules expose the reference's exact state_dict layout and the C-ABI library exports
every symbol include/iggt_b200.h declares."""
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_file_matches_reference_manifest():
    """The product's schema file against the digest of the list extracted from the unmodified reference
    (oracle/make_manifest.py -> oracle/state_manifest.sha256.json)."""
    import hashlib
    rec = json.load(open(os.path.join(ROOT, "oracle", "state_manifest.sha256.json")))
    b = json.load(open(os.path.join(ROOT, "iggt_official_b200", "state_layout.json")))
    assert len(b) == rec["entries"] == 2053
    assert hashlib.sha256(json.dumps(b, separators=(",", ":")).encode()).hexdigest() == rec["sha256_canonical_json"]


@pytest.fixture(scope="module")
def iggt_model():
    from iggt_official_b200.models.vggt import IGGT
    return IGGT()


def test_iggt_state_dict_layout(iggt_model):
    from oracle import weights
    sd = iggt_model.state_dict()
    man = weights.load_manifest()
    assert list(sorted(sd.keys())) == sorted(k for k, _, _ in man)
    for k, shape, dtype in man:
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dtype, k
    assert sum(v.numel() for v in sd.values()) == 1299499573
    # index buffers equal the reference's (restated in oracle.ref_model, checked against the reference
    # by oracle/make_golden.py)
    from oracle import ref_model
    assert torch.equal(sd["part_head.window_cross_attention.relative_position_index_OCA"], ref_model.calculate_rpi_oca(8))
    assert torch.equal(sd["part_head.window_self_atten.relative_position_index_SA"], ref_model.calculate_rpi_sa(8))


def test_load_state_dict_roundtrip_subset(iggt_model):
    from oracle import weights
    sd = weights.make_state_dict(0, "default", prefixes=("camera_head.", "depth_head."))
    missing, unexpected = iggt_model.load_state_dict(sd, strict=False)
    assert not unexpected
    got = iggt_model.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v)


def test_vggt_is_iggt_minus_part():
    from iggt_official_b200.models.vggt import VGGT
    from oracle import weights
    keys = set(VGGT().state_dict().keys())
    want = {k for k, _, _ in weights.load_manifest() if not k.startswith(("part_adaptor.", "part_head."))}
    assert keys == want


def test_forward_rejects_bad_inputs(iggt_model):
    with pytest.raises(RuntimeError):      # no CPU fallback
        iggt_model(torch.rand(2, 3, 28, 28))
    with pytest.raises(ValueError):        # aggregator.py:202-203
        iggt_model.aggregator(torch.rand(1, 2, 4, 28, 28))
    with pytest.raises(AssertionError):    # patch_embed.py:72-73
        iggt_model.aggregator(torch.rand(1, 2, 3, 30, 28))


def test_cabi_exports_every_declared_symbol():
    from iggt_official_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "iggt_b200.h")).read()
    declared = set(re.findall(r"\b(iggt_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("iggt_stream_t")
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in iggt_b200.h but not exported"
    assert declared - {"iggt_version"} == set(_lib.SIGNATURES.keys())


def test_small_attention_view_limit_is_a_clear_error():
    """more views than the camera head's token attention holds in shared memory must raise a descriptive
    error, not a bare launcher status (checked before any device work, so it runs without a GPU)."""
    from iggt_official_b200 import ops
    with pytest.raises(ValueError, match="views"):
        ops.small_attention(torch.zeros(1, 3 * 2048), 1, 300, 16, 128)
