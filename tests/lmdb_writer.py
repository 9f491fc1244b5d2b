"""Test-only writer of an LMDB `data.mdb` (main database, plain byte-string keys), laid out from the description of
LMDB 0.9's on-disk format that vv_records.cpp's reader documents: meta pages 0/1, leaf / branch / overflow pages.
liblmdb is not in this image, so reader and writer share a description rather than the real library (DESIGN.md 5:
the LMDB container is "parity unpinned"; the record VALUES are pinned by the protobuf runtime)."""
import os, struct

P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 1, 2, 4, 8
INVALID = 0xFFFFFFFFFFFFFFFF


def write_lmdb(dirpath, records, psize=4096):
    """records: iterable of (key bytes, value bytes); stored in key order like mdb_put would."""
    records = sorted(records, key=lambda kv: kv[0])
    os.makedirs(dirpath, exist_ok=True)
    pages = {}                          # pgno -> bytes (psize, or a multiple for an overflow run)
    nxt = [2]
    nodemax = (((psize - 16) // 2) & ~1) - 2
    stats = dict(branch=0, leaf=0, overflow=0)

    def alloc(n=1):
        p = nxt[0]; nxt[0] += n
        return p

    def build_page(pgno, flags, nodes):
        body = bytearray(psize)
        upper = psize
        ptrs = []
        for nd in nodes:
            upper -= len(nd) + (len(nd) & 1)
            body[upper:upper + len(nd)] = nd
            ptrs.append(upper)
        lower = 16 + 2 * len(nodes)
        assert lower <= upper, "page overfull"
        body[0:16] = struct.pack("<QHHHH", pgno, 0, flags, lower, upper)
        body[16:lower] = struct.pack("<%dH" % len(ptrs), *ptrs)
        pages[pgno] = bytes(body)

    def flush(level_nodes, flags):
        """pack nodes (first_key, node_bytes) into pages; returns [(first_key, pgno)]"""
        out, cur, used = [], [], 16
        for key, nd in level_nodes:
            need = len(nd) + (len(nd) & 1) + 2
            if cur and used + need > psize:
                out.append(cur); cur, used = [], 16
            cur.append((key, nd)); used += need
        if cur:
            out.append(cur)
        res = []
        for grp in out:
            pgno = alloc()
            nodes = [nd for _, nd in grp]
            if flags == P_BRANCH:       # the first key of a branch page is stored empty
                k0 = grp[0][0]
                nodes[0] = nodes[0][:6] + struct.pack("<H", 0) + nodes[0][8 + len(k0):]
                stats["branch"] += 1
            else:
                stats["leaf"] += 1
            build_page(pgno, flags, nodes)
            res.append((grp[0][0], pgno))
        return res

    leaf_nodes = []
    for key, val in records:
        if 8 + len(key) + len(val) > nodemax:
            n = (16 + len(val) + psize - 1) // psize
            opg = alloc(n)
            blob = bytearray(n * psize)
            blob[0:16] = struct.pack("<QHHI", opg, 0, P_OVERFLOW, n)
            blob[16:16 + len(val)] = val
            pages[opg] = bytes(blob)
            stats["overflow"] += n
            nd = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 1, len(key)) + key + struct.pack("<Q", opg)
        else:
            nd = struct.pack("<HHHH", len(val) & 0xFFFF, len(val) >> 16, 0, len(key)) + key + val
        leaf_nodes.append((key, nd))
    depth, root = 0, INVALID
    if leaf_nodes:
        level = flush(leaf_nodes, P_LEAF); depth = 1
        while len(level) > 1:
            level = flush([(k, struct.pack("<HHHH", p & 0xFFFF, (p >> 16) & 0xFFFF, p >> 32, len(k)) + k) for k, p in level], P_BRANCH)
            depth += 1
        root = level[0][1]
    last_pg = nxt[0] - 1

    def meta(pgno, txnid, live):
        free_db = struct.pack("<IHHQQQQQ", psize, 0, 0, 0, 0, 0, 0, INVALID)
        main_db = (struct.pack("<IHHQQQQQ", 0, 0, depth, stats["branch"], stats["leaf"], stats["overflow"], len(records), root)
                   if live else struct.pack("<IHHQQQQQ", 0, 0, 0, 0, 0, 0, 0, INVALID))
        m = struct.pack("<IIQQ", 0xBEEFC0DE, 1, 0, 1 << 40) + free_db + main_db + struct.pack("<QQ", last_pg if live else 1, txnid)
        body = bytearray(psize)
        body[0:16] = struct.pack("<QHHHH", pgno, 0, P_META, 0, 0)
        body[16:16 + len(m)] = m
        return bytes(body)
    pages[0] = meta(0, 0, False)        # the older, empty transaction: the reader must pick meta page 1
    pages[1] = meta(1, 1, True)
    with open(os.path.join(dirpath, "data.mdb"), "wb") as f:
        p = 0
        while p <= last_pg:
            f.write(pages[p]); p += len(pages[p]) // psize
    open(os.path.join(dirpath, "lock.mdb"), "wb").close()
    return dict(depth=depth, last_pg=last_pg, **stats)
