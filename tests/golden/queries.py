"""Named parity queries. TPC-H texts are the reference's tpch/queries/*.sql statements verbatim
(SQL is data, lowercase as the reference lexer requires); the others exercise the operator
shapes of the reference's test/test_operators.h on TPC-H-shaped tables."""

QUERIES = {
    "q1": """select l_returnflag, l_linestatus, sum(l_quantity) as sum_qty, sum(l_extendedprice) as sum_base_price,
        sum(l_extendedprice * (1 - l_discount)) as sum_disc_price,
        sum(l_extendedprice * (1 - l_discount) * (1 + l_tax)) as sum_charge,
        avg(l_quantity) as avg_qty, avg(l_extendedprice) as avg_price, avg(l_discount) as avg_disc,
        count(*) as count_order
        from lineitem where l_shipdate <= date "1998-9-02"
        group by l_returnflag, l_linestatus order by l_returnflag, l_linestatus""",
    "q6": """select sum(l_extendedprice * l_discount) as revenue from lineitem
        where l_shipdate >= date '1994-01-01' and l_shipdate < date '1995-01-01'
        and l_discount between 0.06 - 0.01 and 0.06 + 0.01 and l_quantity < 24""",
    "q3": """select l_orderkey, sum(l_extendedprice * (1 - l_discount)) as revenue, o_orderdate, o_shippriority
        from customer, orders, lineitem
        where c_mktsegment = 'BUILDING' and c_custkey = o_custkey and l_orderkey = o_orderkey
        and o_orderdate < date '1995-03-15' and l_shipdate > date '1995-03-15'
        group by l_orderkey, o_orderdate, o_shippriority order by revenue desc, o_orderdate limit 10""",
    # the remaining statements of the reference's tpch/queries directory, verbatim
    "q5": """select n_name, sum(l_extendedprice * (1 - l_discount)) as revenue from customer, orders, lineitem, supplier, nation, region where c_custkey = o_custkey and l_orderkey = o_orderkey and l_suppkey = s_suppkey and c_nationkey = s_nationkey and s_nationkey = n_nationkey and n_regionkey = r_regionkey and r_name = 'ASIA' and o_orderdate >= date '1994-01-01' and o_orderdate < date '1995-01-01' group by n_name order by revenue desc""",
    "q10": """select c_custkey, c_name, sum(l_extendedprice * (1 - l_discount)) as revenue, c_acctbal, n_name, c_address, c_phone, c_comment from customer, orders, lineitem, nation where c_custkey = o_custkey and l_orderkey = o_orderkey and o_orderdate >= date '1993-10-01' and o_orderdate < date '1994-01-01' and l_returnflag = 'R' and c_nationkey = n_nationkey group by c_custkey, c_name, c_acctbal, c_phone, n_name, c_address, c_comment order by revenue desc limit 20""",
    "q12": """select l_shipmode, sum(case when o_orderpriority = '1-URGENT' or o_orderpriority = '2-HIGH' then 1 else 0 end) as high_line_count, sum(case when o_orderpriority <> '1-URGENT' and o_orderpriority <> '2-HIGH' then 1 else 0 end) as low_line_count from orders, lineitem where o_orderkey = l_orderkey and l_shipmode in ('MAIL', 'SHIP') and l_commitdate < l_receiptdate and l_shipdate < l_commitdate and l_receiptdate >= date '1994-01-01' and l_receiptdate < date '1995-01-01' group by l_shipmode order by l_shipmode""",
    "q14": """select 100.00 * sum(case when p_type like 'PROMO%' then l_extendedprice * (1 - l_discount) else 0 end) + sum(l_extendedprice * (1 - l_discount)) as promo_revenue from lineitem, part where l_partkey = p_partkey and l_shipdate >= date '1995-09-01' and l_shipdate < date '1995-10-01'""",
    "q19": """select l_extendedprice* (1 - l_discount) from lineitem, part where p_partkey = l_partkey and l_shipinstruct = 'DELIVER IN PERSON' and l_shipmode in ('AIR', 'AIR REG') and ( ( p_brand = 'Brand#12' and p_container in ('SM CASE', 'SM BOX', 'SM PACK', 'SM PKG') and l_quantity >= 1 and l_quantity <= 1 + 10 and p_size between 1 and 5 ) or ( p_brand = 'Brand#23' and p_container in ('MED BAG', 'MED BOX', 'MED PKG', 'MED PACK') and l_quantity >= 10 and l_quantity <= 10 + 10 and p_size between 1 and 10 ) or ( p_brand = 'Brand#34' and p_container in ('LG CASE', 'LG BOX', 'LG PACK', 'LG PKG') and l_quantity >= 20 and l_quantity <= 20 + 10 and p_size between 1 and 15 ) )""",
    # aggregation shapes (test_operators.h: group sum; multi-key sum+count; no groups; groups only)
    "agg_nogroup_minmax": """select min(l_extendedprice), max(l_extendedprice), min(l_shipdate), max(l_shipdate),
        count(*), sum(l_quantity), avg(l_discount) from lineitem where l_quantity < 10""",
    "agg_groups_only": "select l_shipmode from lineitem group by l_shipmode order by l_shipmode",
    "agg_empty": "select count(*), sum(l_quantity) from lineitem where l_quantity < 0",
    "agg_linenumber": """select l_linenumber, count(*) as c, sum(l_extendedprice) as s, min(l_discount) as mn, max(l_tax) as mx
        from lineitem group by l_linenumber order by l_linenumber""",
    "agg_computed_key": """select l_quantity * 2 as q2, sum(l_extendedprice * l_tax) as s from lineitem
        where l_discount > 0.05 group by l_quantity * 2 order by q2""",
    "agg_wrap": """select l_returnflag, sum(l_extendedprice * l_extendedprice * l_extendedprice) as s3
        from lineitem group by l_returnflag order by l_returnflag""",
    "agg_neg_avg": "select avg(c_acctbal), min(c_acctbal), count(*) from customer where c_acctbal < 0.00",
    # selection shapes (decimal lt/gt/or; attr<attr with different scales; date ge/le; and/or mix)
    "sel_or": """select l_orderkey, l_linenumber, l_quantity from lineitem
        where l_quantity > 49 and (l_discount < 0.01 or l_tax > 0.07) and l_orderkey < 2000 order by l_orderkey, l_linenumber""",
    "sel_attr_scales": """select l_orderkey, l_linenumber from lineitem
        where l_quantity < l_tax * 100 and l_orderkey < 3000 order by l_orderkey, l_linenumber""",
    "sel_dates": """select l_orderkey, l_linenumber, l_shipdate, l_commitdate from lineitem
        where l_shipdate >= date '1995-06-01' and l_shipdate <= date '1995-06-30' and l_commitdate < l_receiptdate
        and l_orderkey < 20000 order by l_orderkey, l_linenumber""",
    "sel_neq_char": """select l_returnflag, count(*) as c from lineitem where l_returnflag <> 'N' and l_linestatus = 'F'
        group by l_returnflag order by l_returnflag""",
    "sel_strings": """select c_custkey, c_name, c_mktsegment, c_acctbal from customer
        where c_mktsegment = 'BUILDING  ' and c_acctbal > 9000.00 order by c_custkey""",
    "sel_star_ordered": "select * from customer where c_custkey < 4 order by c_custkey",
    # joins
    "join_orders_lineitem": """select o_orderpriority, count(*) as c, sum(l_extendedprice) as s from orders, lineitem
        where o_orderkey = l_orderkey and l_shipdate > date '1998-06-01' group by o_orderpriority order by o_orderpriority""",
    "join_cust_orders": """select c_mktsegment, count(*) as c, sum(o_totalprice) as s, max(o_orderdate) as d from customer, orders
        where c_custkey = o_custkey and o_orderdate >= date '1998-01-01' group by c_mktsegment order by c_mktsegment""",
    "join_rows": """select o_orderkey, o_orderdate, c_name, c_nationkey from customer, orders
        where c_custkey = o_custkey and o_orderkey < 200 order by o_orderkey""",
    # multi-match hash join (hashjoin.h:118-165; test_operators.h "hash join, duplicates on both sides"):
    # the build side (customer, the smaller table) has many tuples per join key
    "join_dups_rows": """select c_custkey, o_orderkey, o_totalprice from customer, orders
        where c_nationkey = o_custkey order by o_orderkey, c_custkey""",
    "join_dups_agg": """select c_mktsegment, count(*) as c, sum(o_totalprice) as s from customer, orders
        where c_nationkey = o_custkey group by c_mktsegment order by c_mktsegment""",
    "join_dups_both": """select o_orderpriority, count(*) as c, max(c_acctbal) as m from customer, orders
        where c_nationkey = o_shippriority and c_acctbal > 9000.00 and o_orderkey < 3000
        group by o_orderpriority order by o_orderpriority""",
    # three tables in one probe pipeline: customer probes nation (single match), then supplier (multi-match,
    # four suppliers per nation on average), and an attribute of the FIRST build side is still read
    # behind the multi-match probe (hashjoin.h:118-165 emits every match with all joined attributes)
    "join_three_dups": """select n_name, s_name, c_custkey from nation, supplier, customer
        where n_nationkey = c_nationkey and s_nationkey = c_nationkey and c_custkey < 40 order by c_custkey, s_name""",
    "join_three_dups_agg": """select n_name, count(*) as c, sum(s_acctbal) as s, max(c_acctbal) as m from nation, supplier, customer
        where n_nationkey = c_nationkey and s_nationkey = c_nationkey group by n_name order by n_name""",
    # the reference's README microbenchmark (README:69; BASELINE config 5): bigint join + avg + group by
    "micro_join_avg": "select c, avg(d * a) from foo, bar where a = d group by c order by c",
    "micro_join_few_groups": "select c * 0 as g, avg(d * a), count(*), max(a) from foo, bar where a = d group by c * 0",
    # nested-loops join (nestedloopsjoin.h; the planner's fallback for cross products / non-equi joins,
    # planner.h:453-463; test_operators.h has three NLJ shapes incl. a cross product)
    "nlj_cross_filter": """select c_custkey, o_orderkey, c_acctbal, o_totalprice from customer, orders
        where c_custkey < 10 and o_orderkey < 100 and c_acctbal < o_totalprice order by c_custkey, o_orderkey""",
    "nlj_cross_agg": """select count(*) as n, sum(c_acctbal) as s, max(o_totalprice) as m from customer, orders
        where c_custkey < 30 and o_orderkey < 300""",
    "nlj_noneq_group": """select c_mktsegment, count(*) as c, min(o_orderdate) as d from customer, orders
        where c_custkey < 50 and o_orderkey < 400 and c_acctbal * 20 > o_totalprice group by c_mktsegment order by c_mktsegment""",
    "case_sum": """select l_shipmode, sum(case when l_quantity > 25 then 1 else 0 end) as hi,
        sum(case when l_quantity <= 25 then l_extendedprice else 0 end) as lo
        from lineitem group by l_shipmode order by l_shipmode""",
    # ORDER BY beyond one CTA (radix path), multi-key, strings, descending
    "sort_large": """select l_orderkey, l_linenumber, l_extendedprice from lineitem where l_quantity > 40
        order by l_extendedprice desc, l_orderkey, l_linenumber""",
    "sort_strings_small": "select c_name, c_mktsegment, c_acctbal from customer order by c_mktsegment desc, c_name",
    "sort_strings_large": """select o_orderkey, o_orderpriority, o_clerk from orders
        order by o_orderpriority, o_clerk desc, o_orderkey""",
    "agg_string_keys": """select o_orderpriority, o_orderstatus, count(*) as c, sum(o_totalprice) as s from orders
        group by o_orderpriority, o_orderstatus order by o_orderpriority, o_orderstatus""",
    "agg_many_groups": """select l_orderkey, count(*) as c, sum(l_quantity) as q, max(l_shipdate) as d from lineitem
        group by l_orderkey order by l_orderkey""",
    "like_promo": """select count(*) as c from orders where o_comment like '%special%'""",
}
