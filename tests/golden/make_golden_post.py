"""Golden fixtures for ``xpore postprocessing``: run the UNMODIFIED reference filter
(``/root/reference/xpore/scripts/postprocessing.py``) on small hand-made tables and commit its
output.  Runs only in the build container.

    python tests/golden/make_golden_post.py
"""
import importlib.util
import os
import random
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "post")

spec = importlib.util.spec_from_file_location("ref_post", "/root/reference/xpore/scripts/postprocessing.py")
ref_post = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_post)

HEADER = ("id,position,kmer,diff_mod_rate_KO_vs_WT,pval_KO_vs_WT,z_score_KO_vs_WT,mod_rate_KO-rep1,mod_rate_WT-rep1,"
          "coverage_KO-rep1,coverage_WT-rep1,mu_unmod,mu_mod,sigma2_unmod,sigma2_mod,conf_mu_unmod,conf_mu_mod,"
          "mod_assignment,t-test\n")


def table(seed, n, kmers, tie=False):
    rng = random.Random(seed)
    rows = []
    for i in range(n):
        k = rng.choice(kmers)
        d = rng.choice(["higher", "lower", "lower"] if not tie else ["higher", "lower"])
        vals = ["%r" % rng.uniform(-1, 1) for _ in range(13)]
        rows.append(",".join(["ENST%05d" % rng.randrange(50), str(rng.randrange(2000)), k] + vals + [d, "%r" % rng.random()]) + "\n")
    return rows


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "mixed": table(1, 400, ["GGACT", "AGACC", "GGACA", "TGACT", "AAACA", "CCCCC"]),
        "single_direction": [r for r in table(2, 60, ["GGACT", "AGACC"]) if ",lower," in r],
        "empty": [],
    }
    # an exact tie (keeps 'lower', reference line 24)
    tie = table(3, 40, ["GGACT"], tie=True)
    hi = [r for r in tie if ",higher," in r]
    lo = [r for r in tie if ",lower," in r]
    m = min(len(hi), len(lo))
    cases["tie"] = [x for pair in zip(hi[:m], lo[:m]) for x in pair]
    for name, rows in cases.items():
        src = os.path.join(OUT, name + ".diffmod.table")
        with open(src, "w") as f:
            f.write(HEADER)
            f.writelines(rows)
        with tempfile.TemporaryDirectory() as td:
            ref_post.run_postprocessing(src, td)
            with open(os.path.join(td, "majority_direction_kmer_diffmod.table")) as f:
                exp = f.read()
        with open(os.path.join(OUT, name + ".expected.table"), "w") as f:
            f.write(exp)
        print(name, len(rows), "rows ->", exp.count("\n") - 1)


if __name__ == "__main__":
    main()
