"""Extract the known-answer vectors the reference's own tests hold for group arithmetic
on the MSM path into small JSON fixtures (the reference tree does not travel to the GPU box).

Source: /root/reference/ecc/bn254/hash_vectors_test.go:4-112  (hash-to-curve vectors with
explicit affine points; on bn254 G1 (cofactor 1) hash_to_curve gives P = Q0 + Q1, an
absolute known answer for the addition law; G2 points pin the Fp2 tower + twist equation).
        /root/reference/ecc/bls12-381/hash_vectors_test.go (RFC 9380 vectors; on-curve pins)

Run:  python tests/golden/make_golden.py     (in the build container, where /root/reference exists)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse(path):
    src = open(path).read()
    out = {}
    # split by "<name>Vector = "
    for m in re.finditer(r"(\w+Vector)\s*=\s*\w+\{", src):
        name = m.group(1)
        start = m.end()
        nxt = re.search(r"\w+Vector\s*=\s*\w+\{", src[start:])
        body = src[start : start + nxt.start()] if nxt else src[start:]
        cases = []
        for cm in re.finditer(r"msg:\s*\"([^\"]*)\"(.*?)(?=msg:|\Z)", body, re.S):
            case = {"msg": cm.group(1)}
            for pm in re.finditer(r"(\bP|\bQ0|\bQ1|\bQ):\s*point\{\s*\"([^\"]*)\",\s*\"([^\"]*)\",?\s*\}", cm.group(2)):
                case[pm.group(1)] = [pm.group(2), pm.group(3)]
            cases.append(case)
        out[name] = cases
    return out


def main():
    res = {}
    for curve, d in (("bn254", "ecc/bn254"), ("bls12381", "ecc/bls12-381")):
        p = os.path.join(REF, d, "hash_vectors_test.go")
        res[curve] = {"source": d + "/hash_vectors_test.go", "vectors": parse(p)}
    with open(os.path.join(HERE, "hash_vectors.json"), "w") as f:
        json.dump(res, f, indent=1)
    for c in res:
        print(c, {k: len(v) for k, v in res[c]["vectors"].items()})


if __name__ == "__main__":
    main()
