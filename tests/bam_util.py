"""TEST CODE — an independent reader of BGZF / BAM (SAM spec sections 4.1, 4.2) that prints alignment records the way
the reference's SAM lines look (pbsim.cpp:2322-2333), so that BAM output of the engine can be compared with the SAM
text the reference pipes into `samtools view -b`."""
import struct
import zlib

NT16 = "=ACMGRSVTWYHKDBN"


def bgzf_blocks(blob):
    """walk the blocks by their BSIZE fields; returns [(compressed_size, payload)] and checks every header / CRC"""
    out, pos = [], 0
    while pos < len(blob):
        hdr = blob[pos:pos + 18]
        assert hdr[:4] == b"\x1f\x8b\x08\x04", "not a BGZF block at %d" % pos
        xlen, si1, si2, slen, bsize = struct.unpack("<HBBHH", hdr[10:18])
        assert (xlen, si1, si2, slen) == (6, 66, 67, 2)
        size = bsize + 1
        block = blob[pos:pos + size]
        payload = zlib.decompress(block[18:-8], -15)
        crc, isize = struct.unpack("<II", block[-8:])
        assert zlib.crc32(payload) == crc and len(payload) == isize and isize <= 65536
        out.append((size, payload))
        pos += size
    assert pos == len(blob)
    return out


def _int_tag(t, data, p):
    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[t]
    return struct.unpack_from(fmt, data, p)[0], p + struct.calcsize(fmt)


def records_to_sam(raw):
    """BAM alignment records (no file header) -> SAM text"""
    lines, p = [], 0
    while p < len(raw):
        (block_size,) = struct.unpack_from("<I", raw, p)
        rec = raw[p + 4:p + 4 + block_size]
        p += 4 + block_size
        ref_id, pos, l_name, mapq, bin_, n_cigar, flag, l_seq, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", rec, 0)
        assert (ref_id, pos, nref, npos, tlen, n_cigar) == (-1, -1, -1, -1, 0, 0) and bin_ == 4680
        q = 32
        name = rec[q:q + l_name - 1].decode()
        assert rec[q + l_name - 1] == 0
        q += l_name
        packed = rec[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        seq = "".join(NT16[(packed[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        if l_seq & 1:
            assert packed[-1] & 15 == 0
        qual = bytes(x + 33 for x in rec[q:q + l_seq]).decode()
        q += l_seq
        tags = []
        while q < len(rec):
            tag = rec[q:q + 2].decode()
            t = chr(rec[q + 2])
            q += 3
            if t in "cCsSiI":
                v, q = _int_tag(t, rec, q)
                tags.append("%s:i:%d" % (tag, v))
            elif t == "f":
                (v,) = struct.unpack_from("<f", rec, q)
                q += 4
                tags.append("%s:f:%f" % (tag, v))
            elif t == "Z":
                e = rec.index(b"\0", q)
                tags.append("%s:Z:%s" % (tag, rec[q:e].decode()))
                q = e + 1
            elif t == "B":
                st = chr(rec[q])
                (cnt,) = struct.unpack_from("<I", rec, q + 1)
                q += 5
                if st == "f":
                    vals = struct.unpack_from("<%df" % cnt, rec, q)
                    q += 4 * cnt
                    tags.append("%s:B:f,%s" % (tag, ",".join("%.1f" % x for x in vals)))
                else:
                    assert st == "C"
                    vals = rec[q:q + cnt]
                    q += cnt
                    tags.append("%s:B:C%s" % (tag, "".join(",%d" % x for x in vals)))
            else:
                raise AssertionError("unexpected tag type " + t)
        lines.append("\t".join([name, str(flag), "*", "0", str(mapq), "*", "*", "0", "0", seq, qual] + tags) + "\n")
    return "".join(lines).encode()


def parse_bam(blob):
    """whole BAM file -> (header text, SAM records text); checks the magic, n_ref == 0 and the EOF block"""
    blocks = bgzf_blocks(blob)
    assert blocks[-1][1] == b"" and blocks[-1][0] == 28, "missing BGZF end-of-file block"
    raw = b"".join(pl for _, pl in blocks)
    assert raw[:4] == b"BAM\x01"
    (l_text,) = struct.unpack_from("<I", raw, 4)
    text = raw[8:8 + l_text]
    (n_ref,) = struct.unpack_from("<I", raw, 8 + l_text)
    assert n_ref == 0
    return text, records_to_sam(raw[12 + l_text:])
