import sys, json, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from resql_b200 import Engine, Plan
from resql_b200 import tpch_device as TD
from common import load_plan_dict
sf = float(sys.argv[1]) if len(sys.argv)>1 else 10
dev = torch.device('cuda:0')
eng = Engine(0)
orders, li, cust = TD.gen_orders_lineitem(sf, 42, dev, want_orders=False)
n = li['l_quantity'].numel()
print('rows', n)
for q, bpt in (('q6',28),('q1',38)):
    d = load_plan_dict(q)
    names = d['tables'][0]['columns']
    t = eng.upload_device('lineitem', TD.as_device_columns(li, names), n, borrow=True)
    for i in range(5):
        res, tm = eng.execute(Plan(d), {'lineitem': t})
        print(q, 'scan_ms', round(tm.scan_kernel_ms,3), 'Gtuples/s', round(n/tm.scan_kernel_ms/1e6,2), 'GB/s', round(n*bpt/tm.scan_kernel_ms/1e6,1), 'kernels', tm.kernel_launches, 'total kernel ms', round(tm.kernel_ms,3))
    print([c.tolist() for c in res.columns][:3])
    t.free()
# torch cross-check Q6
m = (li['l_shipdate']>=19940101)&(li['l_shipdate']<19950101)&(li['l_discount']>=5)&(li['l_discount']<=7)&(li['l_quantity']<24)
print('torch q6', int((li['l_extendedprice']*li['l_discount'])[m].sum()))
