import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """The oracle's C restatement is built on demand; the product library must already exist
    (python -c 'import __graft_entry__ as g; g.build()') -- tests never build or fall back around it."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    lib = os.path.join(ROOT, "slimt_b200", "libslimt_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "slimt_b200", "csrc")])


@pytest.fixture(scope="session")
def tiny_model(tmp_path_factory):
    """Seeded synthetic tiny11-shaped model in marian binary v1 format (path, parsed items)."""
    from slimt_b200 import synth
    path = str(tmp_path_factory.mktemp("model") / "tiny11.bin")
    synth.write_model(path, synth.make_params(synth.TINY, seed=1234))
    return path, synth.read_model(path)


@pytest.fixture(scope="session")
def eos_model(tmp_path_factory):
    """Same architecture with a large EOS bias so sentences finish at different steps."""
    from slimt_b200 import synth
    path = str(tmp_path_factory.mktemp("model") / "tiny11_eos.bin")
    synth.write_model(path, synth.make_params(synth.TINY, seed=4321, eos_bias=4.8))
    return path, synth.read_model(path)


@pytest.fixture(scope="session")
def shortlist_assets(tmp_path_factory):
    from slimt_b200 import synth
    fr, offs, lists = synth.make_shortlist(vocab=32000, frequent=100, best=100, seed=7)
    path = str(tmp_path_factory.mktemp("sl") / "lex.s2t.bin")
    synth.write_shortlist(path, fr, offs, lists, best=100)
    return path, (fr, offs, lists)


@pytest.fixture(scope="session")
def gpu_ctx():
    from slimt_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
