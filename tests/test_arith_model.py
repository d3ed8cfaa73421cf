"""The word-level algorithms of the device arithmetic (ff63.cuh: mul_b63_w / sqr_b63_w, add_b63_w / sub_b63_w /
neg_b63_w, canon_pm) restated in C with every no-overflow claim and every stated bound of the lazy representation as
an assert (results of one operation are fed into the next), against unsigned __int128 -- runs without a GPU.  The GPU side of the same
check is test_field_ops_lazy_inputs."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_b63_word_model(tmp_path):
    exe = str(tmp_path / "b63_model")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(HERE, "cmodel", "b63_model.c")], check=True)
    out = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ")
