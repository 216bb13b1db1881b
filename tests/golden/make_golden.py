"""Generates the committed golden fixtures. Runs ONLY in the build container (needs
/root/reference and the binaries built by oracle/ref_build/build_ref.sh and
resql_b200/host/build_host.sh):

  * tests/golden/sf001/<name>.out   - output of the REFERENCE ENGINE ITSELF (oracle/_ref/resql-oracle,
    `tofile` format, first line = result schema) on tpch.generate(0.01, seed=42)
  * tests/golden/plans/<name>.json  - the flat plan the host shim lowers from the reference's own
    parser/planner/type-derivation output for the same statement (RESQL_B200_DRY=1)

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from resql_b200 import tpch  # noqa: E402
from queries import QUERIES  # noqa: E402

REF = "/root/reference"
ORACLE = os.path.join(ROOT, "oracle/_ref/resql-oracle")
SHIM = os.path.join(ROOT, "resql_b200/host/resql-b200")


def main():
    out_dir = os.path.join(ROOT, "tests/golden/sf001")
    plan_dir = os.path.join(ROOT, "tests/golden/plans")
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(plan_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        data = tpch.generate(0.01, seed=42)
        loads = ["exec tpch/create.sql", "create table foo ( a bigint, c bigint )", "create table bar ( d bigint )"]
        for name, cols in data.items():
            path = os.path.join(tmp, name + ".bin")
            tpch.to_rows(name, cols).tofile(path)
            loads.append(f"binload {name} {path}")
        for qname, sql in QUERIES.items():
            sql = " ".join(sql.split())
            out = os.path.join(out_dir, qname + ".out")
            r = subprocess.run([ORACLE, "--quiet"] + loads + [f"out {out}", sql], cwd=REF, capture_output=True, text=True)
            ok = "#select" in r.stdout
            env = dict(os.environ, RESQL_B200_DRY="1", RESQL_B200_DUMP_PLAN=os.path.join(plan_dir, qname + ".json"))
            r2 = subprocess.run([SHIM, "--quiet"] + loads + [sql], cwd=REF, capture_output=True, text=True, env=env)
            ok2 = "#select" in r2.stdout
            print(f"{qname:24s} reference={'ok' if ok else 'FAIL'} lowering={'ok' if ok2 else 'FAIL'}")
            if not ok:
                print(r.stdout[-500:], r.stderr[-500:])
            if not ok2:
                print(r2.stdout[-500:], r2.stderr[-500:])
                p = os.path.join(plan_dir, qname + ".json")
                if os.path.exists(p):
                    os.remove(p)


if __name__ == "__main__":
    main()
