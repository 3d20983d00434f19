"""The driver's smoke entry point runs as a GPU test too, so that it cannot rot between rounds."""
import pytest

pytestmark = pytest.mark.gpu


def test_graft_entry_smoke(native_lib):
    import __graft_entry__ as g
    g.smoke()
