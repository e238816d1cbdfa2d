"""The reference's import paths (`ivideogpt.vq_model`, `ivideogpt.transformer`) served by this package (CPU)."""
import importlib
import os
import shutil
import subprocess
import sys

import pytest

from helpers import ROOT


def test_reference_import_paths_resolve_to_the_b200_classes():
    import ivideogpt.transformer as tr
    import ivideogpt.vq_model as vq
    import ivideogpt_b200.transformer as btr
    import ivideogpt_b200.vq_model as bvq
    assert vq.CompressiveVQModel is bvq.CompressiveVQModel and tr.HeadModelWithAction is btr.HeadModelWithAction
    from transformers import AutoModelForCausalLM, LlamaConfig
    from oracle.llama_ref import TINY_LLAMA
    assert isinstance(AutoModelForCausalLM.from_config(LlamaConfig(**TINY_LLAMA)), btr.B200LlamaForCausalLM)
    with pytest.raises(ImportError, match="outside the B200"):
        from ivideogpt.vq_model import Discriminator  # noqa: F401
    with pytest.raises(AttributeError):
        vq.NoSuchThing


def test_alias_serves_the_reference_own_gan_modules_when_installed_next_to_them(tmp_path):
    """INTEGRATION.md install: the alias __init__ replaces the reference's; its discriminator.py / lpips.py stay importable."""
    pkg = tmp_path / "ivideogpt" / "vq_model"
    pkg.mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "ivideogpt", "vq_model", "__init__.py"), pkg / "__init__.py")
    (tmp_path / "ivideogpt" / "__init__.py").write_text("")
    (pkg / "discriminator.py").write_text("class Discriminator:\n    tag = 'reference-own'\n")
    (pkg / "lpips.py").write_text("import a_dependency_that_is_missing\nclass LPIPS:\n    pass\n")
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(1, %r)\n"
            "from ivideogpt.vq_model import Discriminator, CompressiveVQModel\n"
            "assert Discriminator.tag == 'reference-own'\n"
            "try:\n    from ivideogpt.vq_model import LPIPS\nexcept ModuleNotFoundError as e:\n    assert e.name == 'a_dependency_that_is_missing'\n"
            "else:\n    raise SystemExit('expected the missing dependency to surface')\nprint('ok')\n") % (str(tmp_path), ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-800:]
