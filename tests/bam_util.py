"""Tiny independent BAM parser for the tests (gzip handles BGZF's concatenated members)."""
import gzip
import struct


def read_bam(path):
    data = gzip.decompress(open(path, "rb").read())
    assert data[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]
    p += 4
    assert n_ref == 0
    recs = []
    while p < len(data):
        bs = struct.unpack_from("<i", data, p)[0]
        r = data[p + 4:p + 4 + bs]
        p += 4 + bs
        ref, pos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", r, 0)
        q = 32
        name = r[q:q + l_name - 1].decode()
        q += l_name + 4 * n_cig
        packed = r[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(packed[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        qual = list(r[q:q + l_seq])
        q += l_seq
        tags = {}
        while q < len(r):
            tag = r[q:q + 2].decode(); ty = chr(r[q + 2]); q += 3
            if ty == "i":
                tags[tag] = struct.unpack_from("<i", r, q)[0]; q += 4
            elif ty == "f":
                tags[tag] = struct.unpack_from("<f", r, q)[0]; q += 4
            elif ty == "Z":
                e = r.index(b"\x00", q); tags[tag] = r[q:e].decode(); q = e + 1
            elif ty == "B":
                sub = chr(r[q]); n = struct.unpack_from("<i", r, q + 1)[0]; q += 5
                fmt = {"f": "f", "C": "B", "c": "b", "S": "H", "s": "h", "i": "i", "I": "I"}[sub]
                tags[tag] = list(struct.unpack_from("<%d%s" % (n, fmt), r, q)); q += n * struct.calcsize(fmt)
            else:
                raise ValueError("tag type " + ty)
        recs.append(dict(name=name, flag=flag, ref=ref, seq=seq, qual=qual, tags=tags))
    return text, recs
