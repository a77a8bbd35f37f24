"""Minimal BIFF8 (.xls) numeric-cell reader, used only by make_golden.py to read the
reference's datasets/GoogleStock/GOOG.xls (the reference uses `xlrd`, which is not in
this image).  It doubles as an `xlrd` shim: open_workbook(path).sheet_by_index(0).cell_value(r, c).
"""
import struct


def _read_stream(path, name="Workbook"):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\xd0\xcf\x11\xe0\xa1\xb1\x1a\xe1", "not an OLE2 file"
    ssz = 1 << struct.unpack_from("<H", raw, 30)[0]
    n_fat = struct.unpack_from("<I", raw, 44)[0]
    dir_start = struct.unpack_from("<I", raw, 48)[0]
    difat_start, n_difat = struct.unpack_from("<II", raw, 68)

    def sector(i):
        off = 512 + i * ssz
        return raw[off:off + ssz]

    difat = list(struct.unpack_from("<109I", raw, 76))
    nxt = difat_start
    for _ in range(n_difat):
        s = sector(nxt)
        vals = struct.unpack("<%dI" % (ssz // 4), s)
        difat += vals[:-1]
        nxt = vals[-1]
    fat = []
    for s in difat[:n_fat]:
        fat += struct.unpack("<%dI" % (ssz // 4), sector(s))

    def chain(start):
        out, s = [], start
        while s < 0xFFFFFFFC:
            out.append(sector(s))
            s = fat[s]
        return b"".join(out)

    directory = chain(dir_start)
    for i in range(0, len(directory), 128):
        ent = directory[i:i + 128]
        nlen = struct.unpack_from("<H", ent, 64)[0]
        ename = ent[:max(nlen - 2, 0)].decode("utf-16le", "ignore")
        if ename in (name, "Book"):
            start, size = struct.unpack_from("<II", ent, 116)
            assert size >= 4096, "mini-stream workbooks are not supported"
            return chain(start)[:size]
    raise KeyError(name)


def _rk(v):
    if v & 2:
        val = float(struct.unpack("<i", struct.pack("<I", v))[0] >> 2)
    else:
        val = struct.unpack("<d", struct.pack("<II", 0, v & 0xFFFFFFFC))[0]
    return val / 100.0 if v & 1 else val


class _Sheet:
    def __init__(self, cells):
        self._cells = cells

    def cell_value(self, r, c):
        return self._cells[(r, c)]


class _Book:
    def __init__(self, sheets):
        self._sheets = sheets

    def sheet_by_index(self, i):
        return self._sheets[i]


def open_workbook(path):
    data = _read_stream(path)
    pos, sheets, cells, depth = 0, [], None, 0
    n_bof = 0
    while pos + 4 <= len(data):
        rid, rlen = struct.unpack_from("<HH", data, pos)
        body = data[pos + 4:pos + 4 + rlen]
        pos += 4 + rlen
        if rid == 0x0809:            # BOF
            n_bof += 1
            if n_bof > 1:
                cells = {}
        elif rid == 0x000A:          # EOF
            if cells is not None:
                sheets.append(_Sheet(cells))
                cells = None
        elif cells is None:
            continue
        elif rid == 0x0203:          # NUMBER
            r, c, _ = struct.unpack_from("<HHH", body, 0)
            cells[(r, c)] = struct.unpack_from("<d", body, 6)[0]
        elif rid == 0x027E:          # RK
            r, c, _ = struct.unpack_from("<HHH", body, 0)
            cells[(r, c)] = _rk(struct.unpack_from("<I", body, 6)[0])
        elif rid == 0x00BD:          # MULRK
            r, c0 = struct.unpack_from("<HH", body, 0)
            n = (rlen - 6) // 6
            for k in range(n):
                cells[(r, c0 + k)] = _rk(struct.unpack_from("<I", body, 4 + 6 * k + 2)[0])
        elif rid == 0x0006:          # FORMULA with a cached numeric result
            r, c, _ = struct.unpack_from("<HHH", body, 0)
            if body[12:14] != b"\xff\xff":
                cells[(r, c)] = struct.unpack_from("<d", body, 6)[0]
    return _Book(sheets)
