"""Generates tests/golden/like_matrix.json: the accept matrix of the REFERENCE ENGINE's LIKE
(stringLikeCheck, src/qlib/scalar.h:57-120) over a pattern x string grid that covers prefix /
suffix / infix / '_' / overlapping-prefix-and-suffix cases. Runs ONLY in the build container
(needs oracle/_ref/resql-oracle built from /root/reference by oracle/ref_build/build_ref.sh).

    python tests/golden/make_like_matrix.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ORACLE = os.path.join(ROOT, "oracle/_ref/resql-oracle")

STRINGS = ["", "a", "b", "ab", "ba", "aa", "aba", "abab", "abba", "baab", "abc", "abcabc", "aabb", "abcd", "xabcx",
           "special", "a special b", "specialspecial", "PROMO BRUSHED", "PROMO", "PROMOTION", "xPROMO", "sp", "ecial",
           "aXb", "a_b", "a%b", "ab ", " ab", "abab ab", "aaa", "aaaa", "baaab", "ab%ab", "abXab", "ababab", "abcab",
           "cab", "bca", "abcbc"]
PATTERNS = ["", "a", "ab", "%", "%%", "a%", "%a", "%a%", "ab%", "%ab", "%ab%", "ab%ab", "ab%ba", "a%a", "a%b", "a%b%",
            "%a%b", "%a%b%", "a%b%a", "_", "__", "a_", "_b", "a_b", "_%", "%_", "a_%", "%_b", "%a_b%", "ab%%ab", "%ab%ab%",
            "PROMO%", "%special%", "%special%special%", "abc%abc", "abc%bc", "%b%a%", "%aa%", "aa%aa", "%ab_ab%", "a%%", "%%a",
            "ab_", "_ab", "%a_", "_a%", "ab%c", "a%bc", "%abc", "abc%", "%bc%", "b%", "%b", "aaa%a", "a%aaa"]


def main():
    assert len(set(STRINGS)) == len(STRINGS)
    width = 17
    rows = np.zeros((len(STRINGS), width), dtype=np.uint8)
    for i, s in enumerate(STRINGS):
        b = s.encode("latin1")
        rows[i, :len(b)] = list(b)
    match = []
    skipped = []
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "likes.bin")
        rows.tofile(path)
        base = [ORACLE, "--quiet", "create table likes ( s varchar(16) )", f"binload likes {path}"]
        for p in PATTERNS:
            out = os.path.join(tmp, "o.out")
            if os.path.exists(out):
                os.remove(out)
            r = subprocess.run(base + [f"out {out}", f"select s from likes where s like '{p}'"], capture_output=True, text=True)
            if "#select" not in r.stdout:
                # one-character literals are CHAR(1) values in the reference (parseSql.h:104-124), not
                # strings: LIKE on them is not executable there, so they are not part of the matrix
                print(f"{p!r:24s} not executable in the reference (rc={r.returncode}) - skipped")
                skipped.append(p)
                continue
            got = set()
            with open(out, encoding="latin1") as f:
                for line in f.read().split("\n")[1:]:
                    if line.endswith("|"):
                        got.add(line[:-1])
            unknown = got - set(STRINGS)
            assert not unknown, (p, unknown)
            match.append([1 if s in got else 0 for s in STRINGS])
            print(f"{p!r:24s} accepts {sum(match[-1]):2d} of {len(STRINGS)}")
    with open(os.path.join(ROOT, "tests/golden/like_matrix.json"), "w") as f:
        json.dump({"source": "oracle/_ref/resql-oracle (reference stringLikeCheck, qlib/scalar.h:57-120), select s from likes where s like '<pattern>'",
                   "strings": STRINGS, "patterns": [p for p in PATTERNS if p not in skipped], "match": match}, f, indent=0)


if __name__ == "__main__":
    sys.exit(main())
