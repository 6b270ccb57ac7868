"""CPU: the tests' own BGZF / BAM reader (tests/bam_util.py) against a record and a block assembled by hand from the
SAM specification (sections 4.1, 4.2), so that the GPU BAM tests rest on a checked reader."""
import struct
import zlib

from tests import bam_util as B


def _bgzf(payload):
    body = zlib.compress(payload, 6)[2:-4]  # raw deflate
    block = (b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<HBBHH", 6, 66, 67, 2, 18 + len(body) + 8 - 1)
             + body + struct.pack("<II", zlib.crc32(payload), len(payload)))
    return block


def _record():
    name = b"S1/7/2\0"
    seq = "ACGTN"                       # codes 1 2 4 8 15 -> bytes 0x12 0x48 0xF0
    packed = bytes([0x12, 0x48, 0xF0])
    qual = bytes([0, 10, 20, 30, 40])
    tags = (b"cxC\x03" + b"ipBC" + struct.pack("<I", 5) + bytes([9] * 5) + b"npC\x01" + b"pwBC" + struct.pack("<I", 5)
            + bytes([9] * 5) + b"qsC\x00" + b"qeC\x04" + b"rqf" + struct.pack("<f", 0.85)
            + b"snBf" + struct.pack("<I4f", 4, 10.0, 10.0, 10.0, 10.0) + b"zmS" + struct.pack("<H", 300) + b"RGZffffffff\0")
    body = struct.pack("<iiBBHHHIiii", -1, -1, len(name), 255, 4680, 0, 4, len(seq), -1, -1, 0) + name + packed + qual + tags
    return struct.pack("<I", len(body)) + body


def test_record_prints_like_the_reference_sam_line():
    want = ("S1/7/2\t4\t*\t0\t255\t*\t*\t0\t0\tACGTN\t!+5?I\tcx:i:3\tip:B:C,9,9,9,9,9\tnp:i:1\tpw:B:C,9,9,9,9,9\tqs:i:0\t"
            "qe:i:4\trq:f:0.850000\tsn:B:f,10.0,10.0,10.0,10.0\tzm:i:300\tRG:Z:ffffffff\n").encode()
    assert B.records_to_sam(_record() * 2) == want * 2


def test_file_with_header_blocks_and_eof_marker():
    text = b"@HD\tVN:1.5\tSO:unknown\n"
    raw = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", 0) + _record()
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    blob = _bgzf(raw[:20]) + _bgzf(raw[20:]) + eof      # a record may straddle blocks
    hdr, recs = B.parse_bam(blob)
    assert hdr == text
    assert recs.startswith(b"S1/7/2\t4\t*")
    assert [n for n, _ in B.bgzf_blocks(blob)][-1] == 28
