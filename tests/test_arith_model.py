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


def test_block_inverse_tree_model():
    """The three-level product tree of the block-batched inversion (kernels.cu, block_inverse) restated with Python
    integers: K values per thread -> inclusive prefix / suffix scans over the 32 lanes with the kernel's guards ->
    the warps' totals through `scratch`, lane w of warp 0 inverting the product of warps 0..w -> the backward pass.
    Every value must come out as its inverse, zeros as zero (reference: a/0 = 0), whatever the neighbours hold."""
    import random
    p = 9223372036854775783
    rng = random.Random(5)

    def block_inverse(a, K, W):
        # a[warp][lane][k]
        z = [[[v % p == 0 for v in a[w][l]] for l in range(32)] for w in range(W)]
        e = [[[1 if z[w][l][k] else a[w][l][k] for k in range(K)] for l in range(32)] for w in range(W)]
        pre = [[[0] * K for _ in range(32)] for _ in range(W)]
        s = [[0] * 32 for _ in range(W)]
        u = [[0] * 32 for _ in range(W)]
        for w in range(W):
            for l in range(32):
                pre[w][l][0] = e[w][l][0]
                for k in range(1, K):
                    pre[w][l][k] = pre[w][l][k - 1] * e[w][l][k] % p
                s[w][l] = u[w][l] = pre[w][l][K - 1]
            off = 1
            while off < 32:
                sm = [s[w][l] * s[w][l - off] % p if l >= off else None for l in range(32)]      # shfl_up
                um = [u[w][l] * u[w][l + off] % p if l + off < 32 else None for l in range(32)]   # shfl_down
                for l in range(32):
                    if l >= off:
                        s[w][l] = sm[l]
                    if l + off < 32:
                        u[w][l] = um[l]
                off <<= 1
        scratch = [s[w][31] for w in range(W)]
        # warp 0: inclusive scan over the warps' totals, every lane inverts its own prefix
        S = [scratch[l] if l < W else 1 for l in range(32)]
        off = 1
        while off < W:
            sm = [S[l] * S[l - off] % p if l >= off else None for l in range(32)]
            for l in range(32):
                if l >= off:
                    S[l] = sm[l]
            off <<= 1
        IS = [pow(x, p - 2, p) for x in S]
        Sx = [1] + S[:31]
        out = [[[0] * K for _ in range(32)] for _ in range(W)]
        for w in range(W):
            before, inv_upto = Sx[w], IS[w]
            for l in range(32):
                sx = s[w][l - 1] if l > 0 else 1
                ux = u[w][l + 1] if l < 31 else 1
                J = inv_upto * ux % p * (before * sx % p) % p
                for k in range(K - 1, 0, -1):
                    out[w][l][k] = 0 if z[w][l][k] else J * pre[w][l][k - 1] % p
                    J = J * e[w][l][k] % p
                out[w][l][0] = 0 if z[w][l][0] else J
        return out

    for K, W in ((8, 8), (4, 4), (1, 16), (2, 1), (8, 1)):
        a = [[[rng.choice([0, 1, p - 1, p, rng.randrange(p), rng.randrange(p)]) for _ in range(K)] for _ in range(32)]
             for _ in range(W)]
        got = block_inverse(a, K, W)
        for w in range(W):
            for l in range(32):
                for k in range(K):
                    v = a[w][l][k] % p
                    assert got[w][l][k] == (pow(v, p - 2, p) if v else 0), (K, W, w, l, k)


def test_barrett_multiplication_model(tmp_path):
    """ff63::mul_barrett -- the multiplication of the generic-modulus kernels (any odd prime below 2^63, reference:
    FFInt::set_new_prime source/FFInt.cpp:344-353, mod_mul :356-369) -- compiled for the host from the same source the
    device uses, against unsigned __int128 for 265 moduli of every bit length."""
    exe = str(tmp_path / "barrett_model")
    subprocess.run(["g++", "-O2", "-o", exe, os.path.join(HERE, "cmodel", "barrett_model.cpp")], check=True)
    out = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ")
