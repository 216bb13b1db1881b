"""TPC-H-shaped synthetic columns generated directly in HBM (torch is used for device memory
and RNG only). Same distributions as resql_b200.tpch.generate but a different
random stream: parity at these sizes is established through size-independent properties and an
independent torch int64 evaluation of the same query (bench.py / tests), not by row equality."""
import torch


def _ymd(days):
    """days since 1970-01-01 (int64 tensor) -> yyyymmdd (civil_from_days)"""
    z = days + 719468
    era = torch.div(z, 146097, rounding_mode="floor")
    doe = z - era * 146097
    yoe = torch.div(doe - torch.div(doe, 1460, rounding_mode="floor") + torch.div(doe, 36524, rounding_mode="floor")
                    - torch.div(doe, 146096, rounding_mode="floor"), 365, rounding_mode="floor")
    y = yoe + era * 400
    doy = doe - (365 * yoe + torch.div(yoe, 4, rounding_mode="floor") - torch.div(yoe, 100, rounding_mode="floor"))
    mp = torch.div(5 * doy + 2, 153, rounding_mode="floor")
    d = doy - torch.div(153 * mp + 2, 5, rounding_mode="floor") + 1
    m = torch.where(mp < 10, mp + 3, mp - 9)
    y = torch.where(m <= 2, y + 1, y)
    return (y * 10000 + m * 100 + d).to(torch.int32)


D_1992_01_01 = 8035
D_1998_08_02 = 10440
D_1995_06_17 = 9298


def gen_orders_lineitem(sf, seed, device, rank=0, world=1, want_orders=True):
    """Returns (orders, lineitem, customer) dicts of device tensors. lineitem holds the columns
    Q1/Q6/Q3 touch; with world > 1, lineitem is this rank's contiguous row range (row-range
    sharding), orders/customer are replicated (small build sides)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n_orders = max(8, int(round(1_500_000 * sf)))
    n_cust = max(3, int(round(150_000 * sf)))
    n_part = max(1, int(round(200_000 * sf)))
    idx = torch.arange(n_orders, device=device, dtype=torch.int64)
    o_orderkey = (torch.div(idx, 8, rounding_mode="floor") * 32 + idx % 8 + 1).to(torch.int32)
    o_days = torch.randint(D_1992_01_01, D_1998_08_02 + 1, (n_orders,), device=device, generator=g)
    lines = torch.randint(1, 8, (n_orders,), device=device, generator=g)
    cust = torch.randint(0, (n_cust * 2) // 3, (n_orders,), device=device, generator=g)
    o_custkey = (torch.div(cust, 2, rounding_mode="floor") * 3 + cust % 2 + 1).to(torch.int32)   # key % 3 != 0
    orders = None
    if want_orders:
        orders = {"o_orderkey": o_orderkey, "o_custkey": o_custkey, "o_orderdate": _ymd(o_days),
                  "o_shippriority": torch.zeros(n_orders, dtype=torch.int32, device=device)}
    # this rank's order range (contiguous => contiguous lineitem rows)
    lo = n_orders * rank // world
    hi = n_orders * (rank + 1) // world
    lines_r = lines[lo:hi]
    oidx = torch.repeat_interleave(torch.arange(lo, hi, device=device), lines_r)
    n_li = oidx.numel()
    g2 = torch.Generator(device=device)
    g2.manual_seed(seed * 1000003 + rank)
    qty = torch.randint(1, 51, (n_li,), device=device, generator=g2)
    pk = torch.randint(1, n_part + 1, (n_li,), device=device, generator=g2)
    retail = 90000 + (torch.div(pk, 10, rounding_mode="floor") % 20001) + 100 * (pk % 1000)
    ext = qty * retail
    del pk, retail
    disc = torch.randint(0, 11, (n_li,), device=device, generator=g2)
    tax = torch.randint(0, 9, (n_li,), device=device, generator=g2)
    ship = o_days[oidx] + torch.randint(1, 122, (n_li,), device=device, generator=g2)
    receipt = ship + torch.randint(1, 31, (n_li,), device=device, generator=g2)
    ra = torch.where(torch.randint(0, 2, (n_li,), device=device, generator=g2) == 0, ord("R"), ord("A"))
    rf = torch.where(receipt <= D_1995_06_17, ra, ord("N")).to(torch.uint8)
    ls = torch.where(ship > D_1995_06_17, ord("O"), ord("F")).to(torch.uint8)
    del receipt, ra
    lineitem = {"l_orderkey": o_orderkey[oidx], "l_quantity": qty, "l_extendedprice": ext,
                "l_discount": disc, "l_tax": tax, "l_returnflag": rf, "l_linestatus": ls,
                "l_shipdate": _ymd(ship)}
    del ship, oidx
    customer = None
    if want_orders:
        segs = ["AUTOMOBILE", "BUILDING", "FURNITURE", "HOUSEHOLD", "MACHINERY"]
        table = torch.zeros((5, 11), dtype=torch.uint8)
        for i, s in enumerate(segs):
            table[i, :len(s)] = torch.tensor(list(s.encode()), dtype=torch.uint8)
        seg = torch.randint(0, 5, (n_cust,), device=device, generator=g)
        customer = {"c_custkey": torch.arange(1, n_cust + 1, device=device, dtype=torch.int32),
                    "c_mktsegment": table.to(device)[seg].contiguous()}
    return orders, lineitem, customer


RQ_OF = {torch.uint8: (1, 1), torch.int32: (2, 4), torch.int64: (3, 8)}


def as_device_columns(cols, names):
    """dict name -> (ptr, rq_type, width) in `names` order, for Engine.upload_device"""
    out = {}
    for n in names:
        t = cols[n]
        if t.dim() == 2:
            out[n] = (t.data_ptr(), 4, t.shape[1])
        else:
            ty, w = RQ_OF[t.dtype]
            out[n] = (t.data_ptr(), ty, w)
    return out
