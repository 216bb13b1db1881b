"""Canonical text form of result values - restates serializeSqlValue (reference
src/values.h:30-127) and serializeRelation (src/dbdata.h:688-701). This is the parity surface:
the reference's tests compare relations through these strings (test/test_common.h:125-190)."""
import numpy as np

SQL_VARCHAR, SQL_CHAR, SQL_BOOL, SQL_INT, SQL_BIGINT, SQL_DECIMAL, SQL_FLOAT, SQL_DATE = range(8)


def serialize_value(v, sql_type, sql_width):
    if sql_type == SQL_CHAR:
        n = max(int(sql_width), 1)
        if isinstance(v, (bytes, np.bytes_)):
            s = bytes(v).split(b"\0")[0].decode("latin1")
        else:                       # CHAR(1) travels as its payload byte
            s = chr(int(v)) if int(v) != 0 else ""
        return s + " " * max(0, n - len(s))
    if sql_type == SQL_VARCHAR:
        return bytes(v).split(b"\0")[0].decode("latin1")
    if sql_type == SQL_DATE:
        u = int(v) & 0xFFFFFFFF
        return f"{u // 10000}/{u // 100 % 100:02d}/{u % 100:02d}"
    if sql_type in (SQL_INT, SQL_BIGINT):
        return str(int(v))
    if sql_type == SQL_BOOL:
        return "true" if int(v) else "false"
    if sql_type == SQL_DECIMAL:
        scale = int(sql_width) & 0xFF
        x = int(v)
        out = ""
        if x < 0:
            out = "-"
            x = -x
            if x >= 1 << 63:        # INT64_MIN * -1 wraps in the reference
                x -= 1 << 64
        dec = str(x)
        if len(dec) <= scale:
            out += "0." + "0" * (scale - len(dec))
        elif scale > 0:
            dec = dec[: len(dec) - scale] + "." + dec[len(dec) - scale:]
        return out + dec
    raise NotImplementedError(f"serialize for sql type {sql_type}")


def serialize_result(res, sep="|"):
    """Result -> list of lines in the reference's `tofile` format (separator after every field)."""
    lines = []
    for i in range(res.n_rows):
        lines.append("".join(serialize_value(res.columns[c][i], res.sql_types[c], res.sql_widths[c]) + sep
                             for c in range(len(res.columns))))
    return lines
