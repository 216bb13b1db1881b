"""TPC-H-shaped synthetic data, deterministic per seed. The same generator
feeds the reference oracle (as packed tuples in the reference's row format, or `.tbl` text) and
the GPU engine (as columns). Physical types follow tpch/create.sql of the reference:
keys INT, money DECIMAL(12,2) as int64 cents, l_quantity DECIMAL(12,0), flags CHAR(1), dates DATE
(uint32 yyyymmdd), strings NUL-terminated n+1 bytes (src/types.h:213-261)."""
import numpy as np

# name -> ordered [(column, kind, arg)]; kind: int, dec(scale), date, char(n), varchar(n)
SCHEMAS = {
    "lineitem": [("l_orderkey", "int", 0), ("l_partkey", "int", 0), ("l_suppkey", "int", 0),
                 ("l_linenumber", "int", 0), ("l_quantity", "dec", 0), ("l_extendedprice", "dec", 2),
                 ("l_discount", "dec", 2), ("l_tax", "dec", 2), ("l_returnflag", "char", 1),
                 ("l_linestatus", "char", 1), ("l_shipdate", "date", 0), ("l_commitdate", "date", 0),
                 ("l_receiptdate", "date", 0), ("l_shipinstruct", "char", 25), ("l_shipmode", "char", 10),
                 ("l_comment", "varchar", 44)],
    "orders": [("o_orderkey", "int", 0), ("o_custkey", "int", 0), ("o_orderstatus", "char", 1),
               ("o_totalprice", "dec", 2), ("o_orderdate", "date", 0), ("o_orderpriority", "char", 15),
               ("o_clerk", "char", 15), ("o_shippriority", "int", 0), ("o_comment", "varchar", 79)],
    "customer": [("c_custkey", "int", 0), ("c_name", "varchar", 25), ("c_address", "varchar", 40),
                 ("c_nationkey", "int", 0), ("c_phone", "char", 15), ("c_acctbal", "dec", 2),
                 ("c_mktsegment", "char", 10), ("c_comment", "varchar", 117)],
    "part": [("p_partkey", "int", 0), ("p_name", "char", 55), ("p_mfgr", "char", 55), ("p_brand", "char", 10),
             ("p_type", "varchar", 25), ("p_size", "int", 0), ("p_container", "char", 10),
             ("p_retailprice", "dec", 2), ("p_comment", "varchar", 23)],
    "supplier": [("s_suppkey", "int", 0), ("s_name", "char", 25), ("s_address", "varchar", 40), ("s_nationkey", "int", 0),
                 ("s_phone", "char", 15), ("s_acctbal", "dec", 2), ("s_comment", "varchar", 101)],
    "nation": [("n_nationkey", "int", 0), ("n_name", "char", 25), ("n_regionkey", "int", 0), ("n_comment", "varchar", 152)],
    "region": [("r_regionkey", "int", 0), ("r_name", "char", 25), ("r_comment", "varchar", 152)],
    # README microbenchmark of the reference (README:69): select c, avg(d * a) from foo, bar where a = d group by c
    "foo": [("a", "bigint", 0), ("c", "bigint", 0)],
    "bar": [("d", "bigint", 0)],
}

SQL_VARCHAR, SQL_CHAR, SQL_BOOL, SQL_INT, SQL_BIGINT, SQL_DECIMAL, SQL_FLOAT, SQL_DATE = range(8)


def col_dtype(kind, arg):
    if kind == "int" or kind == "date":
        return np.dtype(np.int32)
    if kind == "dec" or kind == "bigint":
        return np.dtype(np.int64)
    if kind == "char" and arg == 1:
        return np.dtype(np.uint8)
    return np.dtype(f"S{arg + 1}")


def tuple_dtype(table):
    """Packed numpy dtype equal to the reference's tuple layout (Schema offsets, schema.h:94-106).
    CHAR(1) occupies 2 bytes (payload + NUL, types.h:232)."""
    names, formats, offsets = [], [], []
    off = 0
    for name, kind, arg in SCHEMAS[table]:
        names.append(name)
        offsets.append(off)
        if kind == "char" and arg == 1:
            formats.append("S2")
            off += 2
        else:
            dt = col_dtype(kind, arg)
            formats.append(dt)
            off += dt.itemsize
    return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": off})


def to_rows(table, cols):
    """columns -> packed tuples (bytes) in the reference row format."""
    dt = tuple_dtype(table)
    n = len(next(iter(cols.values())))
    rows = np.zeros(n, dtype=dt)
    for name, kind, arg in SCHEMAS[table]:
        if kind == "char" and arg == 1:
            rows[name] = cols[name].astype(np.uint8).view("S1")
        else:
            rows[name] = cols[name]
    return rows


def _fmt_dec(v, scale):
    v = int(v)
    s = "-" if v < 0 else ""
    v = abs(v)
    if scale == 0:
        return f"{s}{v}"
    return f"{s}{v // 10 ** scale}.{v % 10 ** scale:0{scale}d}"


def write_tbl(table, cols, path):
    """`.tbl` text the reference's `bulk insert` parses (execute.h:332-388): every field followed
    by '|', decimals with exactly `scale` fraction digits (no rescale on load, expressions.h:430)."""
    n = len(next(iter(cols.values())))
    with open(path, "w") as f:
        for i in range(n):
            parts = []
            for name, kind, arg in SCHEMAS[table]:
                v = cols[name][i]
                if kind == "int":
                    parts.append(str(int(v)))
                elif kind == "dec":
                    parts.append(_fmt_dec(v, arg))
                elif kind == "date":
                    v = int(v)
                    parts.append(f"{v // 10000:04d}-{v // 100 % 100:02d}-{v % 100:02d}")
                elif kind == "char" and arg == 1:
                    parts.append(chr(int(v)))
                else:
                    parts.append(bytes(v).split(b"\0")[0].decode("latin1"))
            f.write("|".join(parts) + "|\n")


_EPOCH = np.datetime64("1970-01-01")


def _ymd(days):
    """days since 1970-01-01 -> yyyymmdd int32"""
    d = _EPOCH + days.astype("timedelta64[D]")
    y = d.astype("datetime64[Y]").astype(np.int64) + 1970
    m = d.astype("datetime64[M]").astype(np.int64) % 12 + 1
    dd = (d - d.astype("datetime64[M]")).astype(np.int64) + 1
    return (y * 10000 + m * 100 + dd).astype(np.int32)


def _days(s):
    return int((np.datetime64(s) - _EPOCH).astype(np.int64))


def _pick(rng, values, n, width):
    arr = np.array([v.encode() for v in values], dtype=f"S{width + 1}")
    return arr[rng.integers(0, len(values), n)]


SEGMENTS = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]
PRIORITIES = ["1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"]
INSTRUCT = ["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"]
MODES = ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"]


def generate_micro(n, groups, seed=42):
    """foo(a bigint, c bigint), bar(d bigint): a and d are permutations of 1..n, c is uniform in
    [0, groups) (columns must be BIGINT - INT has no arithmetic in the reference)."""
    rng = np.random.default_rng(seed * 7919 + 13)
    return {"foo": {"a": (rng.permutation(n) + 1).astype(np.int64), "c": rng.integers(0, groups, n).astype(np.int64)},
            "bar": {"d": (rng.permutation(n) + 1).astype(np.int64)}}


NATIONS = [("ALGERIA", 0), ("ARGENTINA", 1), ("BRAZIL", 1), ("CANADA", 1), ("EGYPT", 4), ("ETHIOPIA", 0), ("FRANCE", 3),
           ("GERMANY", 3), ("INDIA", 2), ("INDONESIA", 2), ("IRAN", 4), ("IRAQ", 4), ("JAPAN", 2), ("JORDAN", 4), ("KENYA", 0),
           ("MOROCCO", 0), ("MOZAMBIQUE", 0), ("PERU", 1), ("CHINA", 2), ("ROMANIA", 3), ("SAUDI ARABIA", 4), ("VIETNAM", 2),
           ("RUSSIA", 3), ("UNITED KINGDOM", 3), ("UNITED STATES", 1)]
REGIONS = ["AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"]
TYPE_1 = ["STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"]
TYPE_2 = ["ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"]
TYPE_3 = ["TIN", "NICKEL", "BRASS", "STEEL", "COPPER"]
CONT_1 = ["SM", "LG", "MED", "JUMBO", "WRAP"]
CONT_2 = ["CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM"]


def _strs(values, width):
    return np.array([v.encode() for v in values], dtype=f"S{width + 1}")


def generate_dims(sf, seed=42):
    """part, supplier, nation, region in the shapes of the TPC-H specification (values drawn uniformly)."""
    rng = np.random.default_rng(seed * 104729 + 7)
    n_part = max(1, int(round(200_000 * sf)))
    n_supp = max(1, int(round(10_000 * sf)))
    pk = np.arange(1, n_part + 1, dtype=np.int64)
    part = {
        "p_partkey": pk.astype(np.int32),
        "p_name": _strs([f"part {i} almond antique" for i in pk], 55),
        "p_mfgr": _strs([f"Manufacturer#{1 + int(i) % 5}" for i in pk], 55),
        "p_brand": _strs([f"Brand#{a}{b}" for a, b in zip(rng.integers(1, 6, n_part), rng.integers(1, 6, n_part))], 10),
        "p_type": _strs([f"{TYPE_1[a]} {TYPE_2[b]} {TYPE_3[c]}" for a, b, c in
                         zip(rng.integers(0, 6, n_part), rng.integers(0, 5, n_part), rng.integers(0, 5, n_part))], 25),
        "p_size": rng.integers(1, 51, n_part).astype(np.int32),
        "p_container": _strs([f"{CONT_1[a]} {CONT_2[b]}" for a, b in zip(rng.integers(0, 5, n_part), rng.integers(0, 8, n_part))], 10),
        "p_retailprice": (90000 + ((pk // 10) % 20001) + 100 * (pk % 1000)).astype(np.int64),
        "p_comment": _pick(rng, ["carefully final", "slyly ironic", "furiously even"], n_part, 23),
    }
    sk = np.arange(1, n_supp + 1, dtype=np.int32)
    snat = rng.integers(0, 25, n_supp).astype(np.int32)
    supplier = {
        "s_suppkey": sk,
        "s_name": _strs([f"Supplier#{i:09d}" for i in sk], 25),
        "s_address": _pick(rng, ["N kD4on9OM Ipw3,gf0J", "89eJ5ksX3ImxJQBvxObC,", "q1,G3Pj6OjIuUYfUoH18BFTKP5aU9bEV3"], n_supp, 40),
        "s_nationkey": snat,
        "s_phone": _strs([f"{10 + int(n)}-{int(k) % 900 + 100}-{int(k) % 9000 + 1000}" for n, k in zip(snat, sk)], 15),
        "s_acctbal": rng.integers(-99999, 1000000, n_supp).astype(np.int64),
        "s_comment": _pick(rng, ["blithely silent requests", "even, bold deposits", "Customer Complaints noted"], n_supp, 101),
    }
    nation = {"n_nationkey": np.arange(25, dtype=np.int32), "n_name": _strs([n for n, _ in NATIONS], 25),
              "n_regionkey": np.array([r for _, r in NATIONS], dtype=np.int32),
              "n_comment": _pick(rng, ["haggle carefully", "final deposits detect", "ironic foxes promise"], 25, 152)}
    region = {"r_regionkey": np.arange(5, dtype=np.int32), "r_name": _strs(REGIONS, 25),
              "r_comment": _pick(rng, ["lar deposits", "hs use ironic requests", "ges thinly even"], 5, 152)}
    return {"part": part, "supplier": supplier, "nation": nation, "region": region}


def generate(sf, seed=42, tables=("lineitem", "orders", "customer", "foo", "bar", "part", "supplier", "nation", "region")):
    """TPC-H-shaped tables at scale factor `sf` (lineitem ~ 6 000 000 * sf rows); foo/bar: the
    microbenchmark tables with 2 000 000 * sf rows and 1000 groups."""
    rng = np.random.default_rng(seed)
    n_orders = max(1, int(round(1_500_000 * sf)))
    n_cust = max(3, int(round(150_000 * sf)))
    n_part = max(1, int(round(200_000 * sf)))
    n_supp = max(1, int(round(10_000 * sf)))
    out = {}

    # ---- orders ---------------------------------------------------------------------------
    o_orderkey = np.arange(1, n_orders + 1, dtype=np.int64)
    o_orderkey = ((o_orderkey - 1) // 8 * 32 + (o_orderkey - 1) % 8 + 1).astype(np.int32)   # dbgen-like sparse keys
    valid_cust = np.arange(1, n_cust + 1)
    valid_cust = valid_cust[valid_cust % 3 != 0]
    o_custkey = valid_cust[rng.integers(0, len(valid_cust), n_orders)].astype(np.int32)
    d0, d1 = _days("1992-01-01"), _days("1998-08-02")
    o_days = rng.integers(d0, d1 + 1, n_orders)
    lines = rng.integers(1, 8, n_orders)
    n_li = int(lines.sum())

    # ---- lineitem -------------------------------------------------------------------------
    oidx = np.repeat(np.arange(n_orders), lines)
    starts = np.cumsum(lines) - lines
    l_linenumber = (np.arange(n_li) - np.repeat(starts, lines) + 1).astype(np.int32)
    l_partkey = rng.integers(1, n_part + 1, n_li).astype(np.int32)
    l_suppkey = rng.integers(1, n_supp + 1, n_li).astype(np.int32)
    qty = rng.integers(1, 51, n_li).astype(np.int64)
    pk = l_partkey.astype(np.int64)
    retail = 90000 + ((pk // 10) % 20001) + 100 * (pk % 1000)
    ext = qty * retail
    disc = rng.integers(0, 11, n_li).astype(np.int64)
    tax = rng.integers(0, 9, n_li).astype(np.int64)
    ship_days = o_days[oidx] + rng.integers(1, 122, n_li)
    commit_days = o_days[oidx] + rng.integers(30, 91, n_li)
    receipt_days = ship_days + rng.integers(1, 31, n_li)
    cutoff = _days("1995-06-17")
    ra = np.where(rng.integers(0, 2, n_li) == 0, ord("R"), ord("A"))
    l_returnflag = np.where(receipt_days <= cutoff, ra, ord("N")).astype(np.uint8)
    l_linestatus = np.where(ship_days > cutoff, ord("O"), ord("F")).astype(np.uint8)
    if "lineitem" in tables:
        out["lineitem"] = {
            "l_orderkey": o_orderkey[oidx], "l_partkey": l_partkey, "l_suppkey": l_suppkey,
            "l_linenumber": l_linenumber, "l_quantity": qty, "l_extendedprice": ext,
            "l_discount": disc, "l_tax": tax, "l_returnflag": l_returnflag,
            "l_linestatus": l_linestatus, "l_shipdate": _ymd(ship_days),
            "l_commitdate": _ymd(commit_days), "l_receiptdate": _ymd(receipt_days),
            "l_shipinstruct": _pick(rng, INSTRUCT, n_li, 25), "l_shipmode": _pick(rng, MODES, n_li, 10),
            "l_comment": _pick(rng, ["regular deposits", "quickly final", "carefully ironic packages", "x"], n_li, 44),
        }
    if "orders" in tables:
        tot = np.zeros(n_orders, dtype=np.int64)
        np.add.at(tot, oidx, ext * (100 - disc) * (100 + tax) // 10000)
        all_f = np.ones(n_orders, dtype=bool)
        all_o = np.ones(n_orders, dtype=bool)
        np.logical_and.at(all_f, oidx, l_linestatus == ord("F"))
        np.logical_and.at(all_o, oidx, l_linestatus == ord("O"))
        status = np.where(all_f, ord("F"), np.where(all_o, ord("O"), ord("P"))).astype(np.uint8)
        clerk = np.array([f"Clerk#{i:09d}".encode() for i in range(1, 1001)], dtype="S16")
        out["orders"] = {
            "o_orderkey": o_orderkey, "o_custkey": o_custkey, "o_orderstatus": status,
            "o_totalprice": tot, "o_orderdate": _ymd(o_days),
            "o_orderpriority": _pick(rng, PRIORITIES, n_orders, 15),
            "o_clerk": clerk[rng.integers(0, 1000, n_orders)],
            "o_shippriority": np.zeros(n_orders, dtype=np.int32),
            "o_comment": _pick(rng, ["furiously special", "pending accounts", "silent asymptotes nag"], n_orders, 79),
        }
    if "customer" in tables:
        ck = np.arange(1, n_cust + 1, dtype=np.int32)
        nat = rng.integers(0, 25, n_cust).astype(np.int32)
        out["customer"] = {
            "c_custkey": ck,
            "c_name": np.array([f"Customer#{i:09d}".encode() for i in ck], dtype="S26"),
            "c_address": _pick(rng, ["IVhzIApeRb ot,c,E", "XSTf4,NCwDVaWNe6tE", "MG9kdTD2WBHm"], n_cust, 40),
            "c_nationkey": nat,
            "c_phone": np.array([f"{10 + int(n)}-{int(k) % 900 + 100}-{int(k) % 9000 + 1000}".encode() for n, k in zip(nat, ck)], dtype="S16"),
            "c_acctbal": rng.integers(-99999, 1000000, n_cust).astype(np.int64),
            "c_mktsegment": _pick(rng, SEGMENTS, n_cust, 10),
            "c_comment": _pick(rng, ["ironic epitaphs nag", "regular platelets", "blithely final"], n_cust, 117),
        }
    if any(t in tables for t in ("part", "supplier", "nation", "region")):
        dims = generate_dims(sf, seed)
        for t in ("part", "supplier", "nation", "region"):
            if t in tables:
                out[t] = dims[t]
    if "foo" in tables or "bar" in tables:
        micro = generate_micro(max(8, int(round(2_000_000 * sf))), 1000, seed)
        for t in ("foo", "bar"):
            if t in tables:
                out[t] = micro[t]
    return out


def sql_type_of(kind, arg):
    """(RQ_SQL_* tag, width) for a schema entry; DECIMAL width = precision<<8 | scale."""
    if kind == "int":
        return SQL_INT, 0
    if kind == "date":
        return SQL_DATE, 0
    if kind == "bigint":
        return SQL_BIGINT, 0
    if kind == "dec":
        return SQL_DECIMAL, (12 << 8) | arg
    if kind == "char":
        return SQL_CHAR, arg
    return SQL_VARCHAR, arg
