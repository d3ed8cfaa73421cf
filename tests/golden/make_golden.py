"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libfirefly_ref.so, built by
`make -C oracle ref` from /root/reference).  Run here, in the authoring container:

    python tests/golden/make_golden.py

Each file: {"source": reference file, "vars": [...], "functions": [text per function, blanks
removed -- what ShuntingYardParser reads between ';'], "rpn": [reference RPN per function],
"cases": [{"prime", "points": [var][point], "values": [function][point]}]}; numbers as decimal strings.
The values come from ShuntingYardParser::evaluate_pre<FFInt> (scalar) of the reference and are
cross-checked against evaluate_pre<FFIntVec<4>> before they are written.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from refs import Ref  # noqa: E402

REF = "/root/reference"
SEED = 0x4D595DF4D0F33173


def splitmix_points(n_vars, n_points, p, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, p - 1, size=(n_vars, n_points), dtype=np.uint64) + np.uint64(1)) % np.uint64(p)


def functions_of(path):
    text = open(path).read()
    out = []
    for chunk in text.split(";"):
        c = chunk.replace(" ", "").replace("\t", "").replace("\n", "")
        if c:
            out.append(c)
    return out


def executable_example(ref, primes):
    """examples/firefly_executable/: the reference's own recorded probes (probes/<prime>.gz, one line per probe
    `zi_order | t | f_1 f_2 f_3`) with the anchor points and shifts they were taken at -- the only numbers the
    reference ships for this path.  The black box behind them is not shipped; the records are those of {0, 1, x1}
    over three variables (f_3 - t is shift[0] of the prime in every line, with the modular wrap on three of them),
    which is what the fixture assumes.  Points as Reconstructor::queue_probes builds them (Reconstructor.hpp:2946-2950):
    x_1 = t + shift_1, x_i = anchor_{i-1}^{zi_{i-1}} * t + shift_i."""
    import gzip
    d = os.path.join(REF, "examples", "firefly_executable")
    anchors = [[int(x) for x in ln.split()] for ln in open(os.path.join(d, "anchor_points")) if ln.strip()]
    shifts = [[int(x) for x in ln.split()] for ln in open(os.path.join(d, "shifts")) if ln.strip()]
    vars_, funs = ["x1", "x2", "x3"], ["0", "1", "x1"]
    doc = {"source": "examples/firefly_executable (anchor_points, shifts, probes/*.gz); black box inferred: {0, 1, x1}",
           "vars": vars_, "functions": funs, "rpn": None, "anchor_points": [[str(x) for x in r] for r in anchors],
           "shifts": [[str(x) for x in r] for r in shifts], "cases": []}
    for pi in range(3):
        p = primes[pi]
        ref.set_prime(p)
        par = ref.parse(funs, vars_)
        if doc["rpn"] is None:
            doc["rpn"] = [par.rpn(i) for i in range(par.n_functions)]
        lines = [ln for ln in gzip.open(os.path.join(d, "probes", f"{pi}.gz"), "rt").read().split("\n") if ln.strip()]
        pts, vals, probes = [], [], []
        for ln in lines:
            zi, t, fv = ln.split("|")
            zi, t, fv = [int(x) for x in zi.split()], int(t), [int(x) for x in fv.split()]
            a = [1] + [pow(anchors[pi][i], zi[i], p) for i in range(2)]
            pts.append([(a[i] * t + shifts[pi][i]) % p for i in range(3)])
            vals.append(fv)
            probes.append({"zi_order": zi, "t": str(t), "line": ln})
        pts = np.array(pts, dtype=np.uint64).T.copy()
        vals = np.array(vals, dtype=np.uint64).T.copy()
        assert np.array_equal(par.eval(pts), vals), "the recorded probes are not those of {0, 1, x1}"
        doc["cases"].append({"prime": str(p), "points": [[str(int(x)) for x in row] for row in pts],
                             "values": [[str(int(x)) for x in row] for row in vals], "probes": probes})
        par.close()
    with open(os.path.join(HERE, "firefly_executable.json"), "w") as f:
        json.dump(doc, f, indent=0, separators=(",", ":"))
    print("firefly_executable: recorded probes per prime", [len(c["probes"]) for c in doc["cases"]])


def main():
    ref = Ref()
    primes = [ref.lib.ffref_prime(i) for i in range(300)]
    executable_example(ref, primes)
    if len(sys.argv) > 1 and sys.argv[1] == "--example-only":
        return
    sets = {
        "s_y_4_v": ("parser_test/s_y_4_v.m", ["x1", "y", "zZ", "W"]),
        "s_y_1_v": ("parser_test/s_y_1_v.m", ["x"]),
        "s_y_safe": ("parser_test/s_y_safe.m", ["x1", "y", "zZ", "W"]),
        "f1": ("benchmarks/f1.m", [f"x{i}" for i in range(1, 21)]),
        "f2": ("benchmarks/f2.m", [f"x{i}" for i in range(1, 6)]),
        "f3": ("benchmarks/f3.m", [f"x{i}" for i in range(1, 6)]),
        "f4": ("benchmarks/f4.m", [f"x{i}" for i in range(1, 6)]),
        "factors1": ("benchmarks/factors1.m", [f"x{i}" for i in range(1, 5)]),
        "factors2": ("benchmarks/factors2.m", [f"x{i}" for i in range(1, 5)]),
        "factors3": ("benchmarks/factors3.m", [f"x{i}" for i in range(1, 21)]),
        "factors4": ("benchmarks/factors4.m", [f"x{i}" for i in range(1, 21)]),
        "factors5": ("benchmarks/factors5.m", [f"x{i}" for i in range(1, 21)]),
    }
    for name, (rel, vars_) in sets.items():
        path = os.path.join(REF, rel)
        doc = {"source": rel, "vars": vars_, "functions": functions_of(path), "rpn": None, "cases": []}
        for pi in (0, 1, 7, 299):
            p = primes[pi]
            ref.set_prime(p)
            par = ref.parse_file(path, vars_)
            if doc["rpn"] is None:
                doc["rpn"] = [par.rpn(i) for i in range(par.n_functions)]
            pts = splitmix_points(len(vars_), 8, p, SEED + pi)
            # edge points: all zero, all one, all p-1, and the survey's documented point
            pts[:, 0] = 0
            pts[:, 1] = 1
            pts[:, 2] = p - 1
            if name == "s_y_4_v":
                pts[:, 3] = np.array([123456789, 9876543210123, 5555555555555555555, 42], dtype=np.uint64)
            vals = par.eval(pts)                       # evaluate_pre<FFInt>
            vals4 = par.eval(pts, bunch=4, threads=1)  # evaluate_pre<FFIntVec<4>>
            assert np.array_equal(vals, vals4)
            doc["cases"].append({
                "prime": str(p),
                "points": [[str(int(x)) for x in row] for row in pts],
                "values": [[str(int(x)) for x in row] for row in vals],
            })
            par.close()
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(doc, f, indent=0, separators=(",", ":"))
        print(name, "functions", len(doc["functions"]), "cases", len(doc["cases"]))

    # hand-written edge cases of the grammar and of the arithmetic (strings mode, no blank stripping)
    vars_ = ["x", "y", "z"]
    funs = [
        "x^(-2)", "-x^2", "(-x)^2", "-x^(-3)", "(-x)^(-3)", "x^0", "0^0", "x/0", "x/(y-y)", "1/x+1/y+1/z",
        "x^y", "x^(-y)", "(x+y)^(z)", "2^(x)", "x^(1/2)", "x^(0-2)", "(x*y)^(-1)/(z^(-1))", "-(x+y)", "+(x-y)",
        "-(-(x))", "((x))", "x+-y", "x-+y", "3/4/5", "x/3/4", "x/(3/4)", "(x+1)/(y+2)/(z+3)", "x^2^3",
        "01234567012345670123", "0x1fffffffffffffffffff", "-99999999999999999999*x", "+123456789012345678901234567890",
        "x - 18446744073709551557", "4294967294", "x^4294967294", "x^(-4294967294)", "x^2147483648", "x^(-2147483647)",
        "y*(x^20+3*x^3*(y-7*z*(1+x)))-z/(x-y)/(x+y)", "(1/x)^(-1)", " x + y * z ", "12abc", "x^(+2)",
        "(x+y+z)^(-2)*(x-y)^(-3)", "x*y/z*x/y", "(x^3-y^3)/(x-y)", "1-1", "x-x", "(x-x)^(-1)", "(x-x)^0",
    ]
    doc = {"source": "hand-written grammar/arithmetic edge cases", "vars": vars_, "functions": funs, "rpn": None,
           "cases": []}
    for pi in (0, 3):
        p = primes[pi]
        ref.set_prime(p)
        par = ref.parse(funs, vars_)
        if doc["rpn"] is None:
            doc["rpn"] = [par.rpn(i) for i in range(par.n_functions)]
        pts = splitmix_points(3, 10, p, SEED + 100 + pi)
        pts[:, 0] = 0
        pts[:, 1] = 1
        pts[:, 2] = p - 1
        pts[:, 3] = np.array([2, 3, 5], dtype=np.uint64)
        pts[:, 4] = np.array([5, 5, 0], dtype=np.uint64)
        pts[:, 5] = np.array([7, p - 3, 2], dtype=np.uint64)       # y = -3: residue exponent >= p/2
        pts[:, 6] = np.array([0, 2, p - 2], dtype=np.uint64)       # zero base
        vals = par.eval(pts)
        doc["cases"].append({
            "prime": str(p),
            "points": [[str(int(x)) for x in row] for row in pts],
            "values": [[str(int(x)) for x in row] for row in vals],
        })
        par.close()
    with open(os.path.join(HERE, "edge_cases.json"), "w") as f:
        json.dump(doc, f, indent=0, separators=(",", ":"))
    print("edge_cases functions", len(funs))


if __name__ == "__main__":
    main()
