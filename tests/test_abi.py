"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/biggan_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_parses_and_library_exports_every_symbol():
    from biggan_tensorflow_b200 import _lib
    assert len(_lib.DECLS) >= 40
    so = ctypes.CDLL(os.path.join(ROOT, "biggan_tensorflow_b200", "libbiggan_b200.so"))
    for name in _lib.DECLS:
        assert hasattr(so, name), "missing export " + name
    assert so.bg_abi_version() == 1


def test_conv_desc_layout_matches_header():
    from biggan_tensorflow_b200._lib import ConvDesc
    assert ctypes.sizeof(ConvDesc) == 16 * 4


def test_geometry_errors_are_reported_without_gpu():
    from biggan_tensorflow_b200 import _lib
    d = _lib.ConvDesc(1, 8, 8, 4, 8, 8, 4, 3, 3, 3, 1, 1, 1, 1, 0, 0)      # stride 3 unsupported
    n = _lib.query("bg_conv2d_wgrad_workspace_bytes", d, 0)
    assert n == 0
    assert b"stride" in _lib.lib.bg_last_error()


def test_option_errors_match_reference_strings():
    """ValueErrors for unknown option strings carry the reference's messages."""
    import pytest
    from biggan_tensorflow_b200 import ops
    with pytest.raises(ValueError, match="Unknown BN type"):
        ops.bn(None, opt={"bn": {"type": "nope"}})
    with pytest.raises(ValueError, match="Unknown BN type"):
        ops.cond_bn(None, None, opt={"bn": {"type": "nope"}})
    with pytest.raises(ValueError, match="Invalid upsampling method"):
        ops.upconv(None, 8, opt={"upsampling_method": "bogus"})
    with pytest.raises(ValueError, match="Invalid generator convolution type"):
        ops.g_conv(None, 8, opt={"g_conv": "bogus"})
    with pytest.raises(ValueError, match="Invalid downsampling method"):
        ops.downconv(None, 8, opt={"downsampling_method": "bogus"})


def test_io_library_exports_every_symbol_of_its_header():
    """include/biggan_io.h (host-side input decoding) <-> libbiggan_io.so."""
    import re
    src = open(os.path.join(ROOT, "include", "biggan_io.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(bgio_\w+)\s*\(", src)
    assert set(names) == {"bgio_last_error", "bgio_png_info", "bgio_png_decode"}
    so = ctypes.CDLL(os.path.join(ROOT, "biggan_tensorflow_b200", "libbiggan_io.so"))
    for n in names:
        assert hasattr(so, n), "missing export " + n


def test_option_errors_of_the_config_level_switches():
    """BigGAN.py:71-84 (activation), :257-271 (regularisation), utils.py:204 (regulariser type)."""
    import pytest
    from biggan_tensorflow_b200 import model, utils
    with pytest.raises(ValueError, match="Unknown activation function: swish"):
        model._opts(model.Config(activation="swish"), True)
    with pytest.raises(ValueError, match="Unknown regularization method: l3"):
        model._opts(model.Config(g_regularization="l3"), True)
    g_opt, d_opt = model._opts(model.Config(g_regularization="ortho_cosine"), is_training=False)     # sampling graph: no regulariser
    assert g_opt["conv"]["regularizer"] is None and g_opt["fc_regularizer"] is None
    g_opt, d_opt = model._opts(model.Config(g_regularization="ortho", activation="lrelu", bn_in_d=True, downsampling_method="pool_only"), True)
    assert g_opt["conv"]["regularizer"].type == "ortho" and "regularizer" in g_opt["conv"] and d_opt["conv"]["regularizer"] is None
    assert d_opt["bn_in_d"] is True and d_opt["downsampling_method"] == "pool_only" and g_opt["act"] is d_opt["act"]
    with pytest.raises(ValueError, match="Unknown regularization method."):
        utils.orthogonal_regularizer(1e-4, type="nope")._kind()
