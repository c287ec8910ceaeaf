"""CPU oracle: a numpy restatement of the reference's scan-filter-aggregate path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this module, and only as the checker.  The product (eventql_b200/, include/) never
imports it; the product path fails loudly when its CUDA library is missing.

Parity status: PINNED.  tests/test_oracle_pinning.py checks this restatement against
  * the reference's golden files test/sql/00001 (all 213 `time` values) and 00002 (count = 213),
    copied to tests/golden/ together with the fixture test/sql_testdata/testtbl.cst,
  * the known answers written down in io/cstable/cstable_test.cc:74-240,483-521,587-755,
  * outputs of the reference itself (oracle/_ref/evqlref, built from /root/reference by
    oracle/build_ref.py) committed as tests/golden/ref_*.json by tests/golden/make_golden*.py: query rows
    (ref_results.json), ORDER BY / LIMIT rows in order (ref_orderby.json), PartialGroupByExpression rows with the
    reference's own .qc cache entries and result frames (ref_partial.json), string columns and string queries on
    reference-written tables (ref_strings.json, tests/test_strings_lsm.py), and the rows of the reference's own
    eventql::PartitionCursor over LSM partitions for the visibility filter (ref_lsm.json).

Everything below follows the reference file:line cited next to it (paths relative to
/root/reference/src/eventql/).  Integer work is bit-exact (numpy uint64/int64 wrap like C);
float64 sums are accumulated in a different order than the reference, hence the 1e-9 relative
tolerance stated by BASELINE.json.
"""
from __future__ import annotations

import hashlib
import os
import struct
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventql_b200 import plan as P  # pure-Python plan description shared with the C-ABI binding  # noqa: E402

MAGIC = b"\x23\x17\x23\x17"          # io/cstable/cstable.h:137
PAGE_SIZE = 512 * 1024               # columns/page_writer_*.h:34
BITPACK_PAGE_VALUES = 1024 * 128     # page_writer_bitpacked.cc:46-53 (16 * maxbits * 1024 bytes)
SECTOR = 512                         # cstable.h:139-143
META_POS, META_SIZE = 14, 48         # cstable.h:218-219


class OracleError(Exception):
    pass


# ------------------------------------------------------------------------------------------------
# primitive codecs
# ------------------------------------------------------------------------------------------------

def bits(v: int) -> int:
    """libsimdcomp bits(): 32 - clz(v), 0 for v == 0 (simdcomputil.c)."""
    return int(v).bit_length()


def read_varuint(buf: bytes, pos: int) -> Tuple[int, int]:
    v = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not (b & 0x80):
            return v, pos
        shift += 7


def write_varuint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def leb128_decode(stream: np.ndarray, n: int) -> np.ndarray:
    """columns/page_reader_leb128.cc:50-72 - unsigned LEB128, pages treated as one byte stream."""
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    stream = np.asarray(stream, dtype=np.uint8)
    ends = np.flatnonzero((stream & 0x80) == 0)
    if len(ends) < n:
        raise OracleError("end of column reached")
    ends = ends[:n].astype(np.int64)
    starts = np.empty(n, dtype=np.int64)
    starts[0] = 0
    starts[1:] = ends[:-1] + 1
    lens = ends - starts + 1
    vals = np.zeros(n, dtype=np.uint64)
    for k in range(int(lens.max())):
        m = lens > k
        if k >= 10:
            # (b & 0x7f) << (7*k) with 7*k >= 64 is undefined in the reference; such encodings are not produced
            raise OracleError("over-long LEB128 value")
        byte = stream[starts[m] + k].astype(np.uint64) & np.uint64(0x7F)
        vals[m] |= byte << np.uint64(7 * k)
    return vals


def leb128_encode(values: np.ndarray) -> np.ndarray:
    """columns/page_writer_leb128.cc:38-66."""
    v = np.asarray(values, dtype=np.uint64)
    n = len(v)
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    nbits = np.zeros(n, dtype=np.int64)
    tmp = v.copy()
    # bit length via repeated shifts (vectorised)
    for s in (32, 16, 8, 4, 2, 1):
        m = tmp >= (np.uint64(1) << np.uint64(s))
        nbits[m] += s
        tmp[m] >>= np.uint64(s)
    nbits += (tmp > 0).astype(np.int64)
    lens = np.maximum(1, (nbits + 6) // 7)
    offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    out = np.zeros(int(offs[-1]), dtype=np.uint8)
    for k in range(int(lens.max())):
        m = lens > k
        byte = ((v[m] >> np.uint64(7 * k)) & np.uint64(0x7F)).astype(np.uint8)
        cont = (lens[m] > k + 1)
        byte |= (cont.astype(np.uint8) << 7)
        out[offs[:-1][m] + k] = byte
    return out


def bitunpack_vertical(stream: np.ndarray, n: int, b: int) -> np.ndarray:
    """libsimdcomp simdunpack layout (deps/3rdparty/libsimdcomp/simdbitpacking.c:13793):
    block of 128 values = b 128-bit words; value i -> lane i%4, sequence i//4, bit offset (i//4)*b
    inside the lane's private bit stream."""
    if b == 0 or n == 0:
        return np.zeros(n, dtype=np.uint64)
    words = np.frombuffer(np.ascontiguousarray(stream, dtype=np.uint8).tobytes(), dtype="<u4").astype(np.uint64)
    i = np.arange(n, dtype=np.int64)
    block = i // 128
    w = i % 128
    lane = w % 4
    seq = w // 4
    o = seq * b
    idx = block * (4 * b) + 4 * (o // 32) + lane
    shift = (o % 32).astype(np.uint64)
    if idx.max() >= len(words):
        raise OracleError("bit-packed stream too short")
    val = words[idx] >> shift
    spill = (o % 32) + b > 32
    if spill.any():
        idx2 = np.minimum(idx + 4, len(words) - 1)
        hi = words[idx2] << (np.uint64(32) - shift)
        val = np.where(spill, val | hi, val)
    mask = np.uint64((1 << b) - 1)
    return val & mask


def bitpack_vertical(values: np.ndarray, b: int) -> np.ndarray:
    """simdpackwithoutmask (simdbitpacking.c:13868), values must fit b bits; last block zero-filled
    (columns/page_writer_bitpacked.cc:63-80)."""
    n = len(values)
    if b == 0 or n == 0:
        return np.zeros(0, dtype=np.uint8)
    nblocks = (n + 127) // 128
    v = np.zeros(nblocks * 128, dtype=np.uint64)
    v[:n] = np.asarray(values, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    words = np.zeros(nblocks * 4 * b + 4, dtype=np.uint64)
    i = np.arange(nblocks * 128, dtype=np.int64)
    block = i // 128
    w = i % 128
    lane = w % 4
    seq = w // 4
    o = seq * b
    idx = block * (4 * b) + 4 * (o // 32) + lane
    shift = (o % 32).astype(np.uint64)
    lo = (v << shift) & np.uint64(0xFFFFFFFF)
    np.bitwise_or.at(words, idx, lo)
    spill = (o % 32) + b > 32
    hi = v >> (np.uint64(32) - shift)
    np.bitwise_or.at(words, idx[spill] + 4, hi[spill])
    return words[: nblocks * 4 * b].astype("<u4").view(np.uint8)


# ------------------------------------------------------------------------------------------------
# cstable reader (io/cstable/cstable.cc, cstable_reader.cc, columns/*)
# ------------------------------------------------------------------------------------------------

@dataclass
class ColumnConfig:            # io/cstable/cstable.h:148-157
    name: str
    column_id: int
    logical_type: int
    storage_type: int
    rlevel_max: int
    dlevel_max: int
    body_offset: int = 0       # v0.1.0 only
    body_size: int = 0


@dataclass
class CSTableFile:
    version: int
    num_rows: int
    columns: Dict[str, ColumnConfig]
    data: bytes
    index: List[Tuple[int, int, int, int]] = field(default_factory=list)   # (entry_type, column_id, offset, size)

    def pages(self, column_id: int, kind: int) -> List[Tuple[int, int]]:
        """PageManager::getPages (page_manager.cc:158-171): pages of one stream in index order."""
        return [(o, s) for (t, c, o, s) in self.index if t == kind and c == column_id]

    def stream(self, col: ColumnConfig, kind: int) -> Tuple[np.ndarray, int]:
        """Logical stream = concatenation of the pages; bit-packed streams lose their 4-byte max_value
        header, returned separately (page_reader_bitpacked.cc:30-49)."""
        assert self.version == 2
        pages = self.pages(col.column_id, kind)
        bitpacked = kind != P.STREAM_DATA or col.storage_type in (P.ENC_UINT32_BITPACKED, P.ENC_BOOLEAN_BITPACKED)
        parts = []
        maxv = 0
        for k, (o, s) in enumerate(pages):
            raw = self.data[o:o + s]
            if len(raw) != s:
                raise OracleError("read() failed")
            if bitpacked and k == 0:
                maxv = struct.unpack_from("<I", raw, 0)[0]
                raw = raw[4:]
            parts.append(np.frombuffer(raw, dtype=np.uint8))
        arr = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
        return arr, maxv


def parse_cstable(data: bytes) -> CSTableFile:
    """cstable::readHeader (io/cstable/cstable.cc:35-84) + v0_1_0::readHeader (:89-132) +
    v0_2_0::readHeader/readMetaBlock/readIndex (:152-171, :200-255)."""
    if data[:4] != MAGIC:
        raise OracleError("not a valid cstable file")
    vnum = data[4]
    if vnum == 1:
        pos = 6
        _flags, num_rows, ncols = struct.unpack_from("<QQI", data, pos)
        pos += 20
        cols = {}
        for _ in range(ncols):
            storage, nlen = struct.unpack_from("<II", data, pos)
            pos += 8
            name = data[pos:pos + nlen].decode()
            pos += nlen
            rmax, dmax, boff, bsize = struct.unpack_from("<IIQQ", data, pos)
            pos += 24
            logical = {P.ENC_BOOLEAN_BITPACKED: P.COL_BOOLEAN, P.ENC_FLOAT_IEEE754: P.COL_FLOAT,
                       P.ENC_STRING_PLAIN: P.COL_STRING}.get(storage, P.COL_UNSIGNED_INT)
            cols[name] = ColumnConfig(name, 0, logical, storage, rmax, dmax, boff, bsize)
        return CSTableFile(1, num_rows, cols, data)
    if vnum != 2:
        raise OracleError("unsupported cstable version: %d" % vnum)
    metas = []
    for k in range(2):
        mb = data[META_POS + k * META_SIZE: META_POS + (k + 1) * META_SIZE]
        if hashlib.sha1(mb[:28]).digest() == mb[28:48]:
            metas.append(struct.unpack_from("<QQQI", mb, 0))
    if not metas:
        raise OracleError("can't open cstable: no valid metablocks found")
    # cstable.cc:64-76: the block with the larger transaction id wins
    txid, num_rows, index_offset, index_size = max(metas, key=lambda m: m[0])
    pos = META_POS + 2 * META_SIZE + 128
    ncols, pos = read_varuint(data, pos)
    cols = {}
    for _ in range(ncols):
        logical, pos = read_varuint(data, pos)
        storage, pos = read_varuint(data, pos)
        cid, pos = read_varuint(data, pos)
        nlen, pos = read_varuint(data, pos)
        name = data[pos:pos + nlen].decode()
        pos += nlen
        rmax, pos = read_varuint(data, pos)
        dmax, pos = read_varuint(data, pos)
        cols[name] = ColumnConfig(name, cid, logical, storage, rmax, dmax)
    index = []
    pos = index_offset
    n, pos = read_varuint(data, pos)
    for _ in range(n):
        t, pos = read_varuint(data, pos)
        c, pos = read_varuint(data, pos)
        o, pos = read_varuint(data, pos)
        s, pos = read_varuint(data, pos)
        index.append((t, c, o, s))
    return CSTableFile(2, num_rows, cols, data, index)


def read_cstable(path: str) -> CSTableFile:
    with open(path, "rb") as f:
        return parse_cstable(f.read())


@dataclass
class DecodedColumn:
    values: np.ndarray      # uint64 raw bits (float columns: IEEE-754 bits), one per record, 0 where NULL
    present: np.ndarray     # bool, one per record
    sql_type: int


def sql_type_of(col: ColumnConfig) -> int:
    """sql/CSTableScanProvider.cc:79-107: cstable logical type -> SType (DATETIME -> UINT64)."""
    return {P.COL_BOOLEAN: P.BOOL, P.COL_UNSIGNED_INT: P.UINT64, P.COL_SIGNED_INT: P.INT64,
            P.COL_FLOAT: P.FLOAT64, P.COL_STRING: P.STRING, P.COL_DATETIME: P.UINT64}[col.logical_type]


def _decode_data(storage: int, stream: np.ndarray, n: int, maxv: int) -> np.ndarray:
    if storage == P.ENC_UINT64_LEB128:
        return leb128_decode(stream, n)
    if storage in (P.ENC_UINT64_PLAIN, P.ENC_FLOAT_IEEE754):      # page_reader_uint64.cc:50-70, page_reader_ieee754.cc:38-59
        if len(stream) < 8 * n:
            raise OracleError("end of column reached")
        return np.frombuffer(stream[:8 * n].tobytes(), dtype="<u8").astype(np.uint64)
    if storage == P.ENC_UINT32_PLAIN:                              # page_reader_uint32.cc:50-71
        if len(stream) < 4 * n:
            raise OracleError("end of column reached")
        return np.frombuffer(stream[:4 * n].tobytes(), dtype="<u4").astype(np.uint64)
    if storage in (P.ENC_UINT32_BITPACKED, P.ENC_BOOLEAN_BITPACKED):
        return bitunpack_vertical(stream, n, bits(maxv))
    raise OracleError("unsupported storage type %d" % storage)


def decode_column(f: CSTableFile, name: str) -> DecodedColumn:
    """One value per record, the way FastCSTableScan consumes a flat column
    (sql/CSTableScan.cc:860-968 over columns/column_reader_uint.cc:92-115 / v1 readers):
    value is read from the DATA stream only if dlevel == dlevel_max."""
    col = f.columns[name]
    if col.rlevel_max > 0:
        raise OracleError("repeated column %s: FastCSTableScan reads one entry per record (unsupported)" % name)
    if col.logical_type in (P.COL_STRING, P.COL_SUBRECORD, P.COL_SIGNED_INT):
        raise OracleError("column type of %s is outside the numeric scan path" % name)
    n = f.num_rows
    if f.version == 2:
        if col.dlevel_max > 0:
            dstream, dmaxv = f.stream(col, P.STREAM_DLEVEL)
            dl = bitunpack_vertical(dstream, n, bits(dmaxv))
            present = dl == np.uint64(col.dlevel_max)
        else:
            present = np.ones(n, dtype=bool)
        stream, maxv = f.stream(col, P.STREAM_DATA)
        nn = int(present.sum())
        vals = _decode_data(col.storage_type, stream, nn, maxv)
    else:
        # columns/v1/ColumnReader.h:37-55: u64 num_vals | rlvl_size | dlvl_size | data_size | rlvl | dlvl | data
        body = f.data[col.body_offset: col.body_offset + col.body_size]
        num_vals, rsize, dsize, datasize = struct.unpack_from("<QQQQ", body, 0)
        dl_stream = np.frombuffer(body[32 + rsize: 32 + rsize + dsize], dtype=np.uint8)
        data = np.frombuffer(body[32 + rsize + dsize: 32 + rsize + dsize + datasize], dtype=np.uint8)
        dl = bitunpack_vertical(dl_stream, n, bits(col.dlevel_max))     # util/util/BitPackDecoder.cc:29-49
        present = dl == np.uint64(col.dlevel_max)
        nn = int(present.sum())
        maxv = 0
        st = col.storage_type
        if st == P.ENC_UINT32_BITPACKED:          # columns/v1/BitPackedIntColumnReader.cc:31-43
            maxv = struct.unpack_from("<I", data.tobytes(), 0)[0] if len(data) >= 4 else 0
            data = data[4:]
        elif st == P.ENC_BOOLEAN_BITPACKED:       # columns/v1/BooleanColumnReader.cc:31-38 (max_val = 1)
            maxv = 1
        vals = _decode_data(st, data, nn, maxv)
    out = np.zeros(n, dtype=np.uint64)
    out[present] = vals
    st = sql_type_of(col)
    if st == P.BOOL:
        out = (out > 0).astype(np.uint64)       # column_reader_uint.cc:76-90 readBoolean: value > 0
    return DecodedColumn(out, present, st)


def decode_string_column(f: CSTableFile, name: str) -> List[Optional[bytes]]:
    """One string (or None for NULL) per record of a flat STRING_PLAIN column, the way FastCSTableScan::fetchColumnString
    consumes it (sql/CSTableScan.cc:970-995): StringColumnReader::readString reads a value only where the definition level
    equals dlevel_max (columns/column_reader_string.cc).  v0.2.0 values are `varuint length + bytes` and may straddle pages
    (columns/page_reader_lenencstring.cc:37-62); v0.1.0 values are `u32 length + bytes` (columns/v1/StringColumnReader.cc:94-113)."""
    col = f.columns[name]
    if col.storage_type != P.ENC_STRING_PLAIN:
        raise OracleError("invalid storage type for string column '%s'" % name)
    if col.rlevel_max > 0:
        raise OracleError("repeated column %s: FastCSTableScan reads one entry per record (unsupported)" % name)
    n = f.num_rows
    if f.version == 2:
        if col.dlevel_max > 0:
            dstream, dmaxv = f.stream(col, P.STREAM_DLEVEL)
            present = bitunpack_vertical(dstream, n, bits(dmaxv)) == np.uint64(col.dlevel_max)
        else:
            present = np.ones(n, dtype=bool)
        data = f.stream(col, P.STREAM_DATA)[0].tobytes()
    else:
        body = f.data[col.body_offset: col.body_offset + col.body_size]
        _num_vals, rsize, dsize, datasize = struct.unpack_from("<QQQQ", body, 0)
        dl_stream = np.frombuffer(body[32 + rsize: 32 + rsize + dsize], dtype=np.uint8)
        present = bitunpack_vertical(dl_stream, n, bits(col.dlevel_max)) == np.uint64(col.dlevel_max)
        data = body[32 + rsize + dsize: 32 + rsize + dsize + datasize]
    out: List[Optional[bytes]] = []
    pos = 0
    for i in range(n):
        if not present[i]:
            out.append(None)
            continue
        if f.version == 2:
            if pos >= len(data):
                raise OracleError("end of column reached")
            ln, pos = read_varuint(data, pos)
        else:
            if pos + 4 > len(data):
                raise OracleError("end of column reached")
            ln = struct.unpack_from("<I", data, pos)[0]
            pos += 4
        if pos + ln > len(data):
            raise OracleError("end of column reached")
        out.append(data[pos:pos + ln])
        pos += ln
    return out


def pack_svector_strings(vals: Sequence[Optional[bytes]]) -> bytes:
    """The packed STRING SVector FastCSTableScan::fetchColumnString builds (sql/CSTableScan.cc:970-995, svalue.cc:533-549):
    present -> [u32 length][bytes][tag 0]; NULL -> [u32 0][tag STAG_NULL]."""
    out = bytearray()
    for v in vals:
        if v is None:
            out += struct.pack("<I", 0) + bytes([P.STAG_NULL])
        else:
            out += struct.pack("<I", len(v)) + v + b"\0"
    return bytes(out)


def unpack_svector_strings(buf: bytes, n: int) -> List[Optional[bytes]]:
    out: List[Optional[bytes]] = []
    pos = 0
    for _ in range(n):
        ln = struct.unpack_from("<I", buf, pos)[0]
        tag = buf[pos + 4 + ln]
        out.append(None if (tag & P.STAG_NULL) else bytes(buf[pos + 4: pos + 4 + ln]))
        pos += 5 + ln
    if pos != len(buf):
        raise OracleError("trailing bytes in a string SVector")
    return out


@dataclass
class LsmSegment:
    """One table PartitionCursor::openNextTable visits (server/sql/partition_cursor.cc:82-155), in its order: head arena,
    compacting arena, then the on-disk LSM tables newest first."""
    table: CSTableFile
    skiplist: Optional[np.ndarray] = None   # arena skiplist (PartitionArena::SkiplistReader), one bool per row
    use_skip_column: bool = False           # tbl->has_skiplist(): read the __lsm_skip column
    needs_filter: Optional[bool] = True     # None: the cursor's rule for on-disk tables (partition_cursor.cc:149-155)
    has_updates: bool = False               # LSMTableRef::has_updates
    oldest: bool = False                    # tblidx == 0


def lsm_visibility(segments: Sequence[LsmSegment]) -> List[Optional[np.ndarray]]:
    """The row filters PartitionCursor::openNextTable hands to setFilter (server/sql/partition_cursor.cc:157-194, :216-218):
    a row is dropped if it is skipped or if a row seen earlier (in an earlier table, or earlier in this one) was an update
    with the same __lsm_id; visible update rows add their id to the set.  None = no filter (needs_filter == false: the
    table's ids are not recorded either).  Pinned: tests/golden/ref_lsm.json holds the rows the reference's own
    PartitionCursor returns over partitions of on-disk tables (oracle/ref_tools/evqlref.cc `sql -S`); only the arena
    skiplist override has no reference run behind it."""
    id_set = set()
    out: List[Optional[np.ndarray]] = []
    for seg in segments:
        n = seg.table.num_rows
        needs = seg.needs_filter
        if needs is None:                                # partition_cursor.cc:139-155
            needs = True
            has_skiplist = seg.use_skip_column or seg.skiplist is not None
            if not has_skiplist and seg.oldest and not id_set:
                needs = False
            if not has_skiplist and not seg.has_updates and not id_set:
                needs = False
        if not needs:
            out.append(None)
            continue
        ids = decode_string_column(seg.table, "__lsm_id")
        upd = decode_column(seg.table, "__lsm_is_update").values != 0
        skipcol = decode_column(seg.table, "__lsm_skip").values != 0 if seg.use_skip_column else np.zeros(n, dtype=bool)
        flt = np.ones(n, dtype=bool)
        for i in range(n):
            ident = ids[i] if ids[i] is not None else b""
            if len(ident) != 20:                         # SHA1Hash(const void*, size_t): util/SHA1.cc:79-85
                raise OracleError("invalid SHA1Hash")
            skip = bool(skipcol[i])
            if seg.skiplist is not None:
                skip = bool(seg.skiplist[i])
            if skip or ident in id_set:
                flt[i] = False
            elif upd[i]:
                id_set.add(ident)
        out.append(flt)
    return out


# ------------------------------------------------------------------------------------------------
# cstable writer, v0.2.0 (io/cstable/cstable_writer.cc:267-293, cstable.cc:138-198, page_manager.cc:45-75)
# ------------------------------------------------------------------------------------------------

@dataclass
class WriteColumn:
    name: str
    logical_type: int
    encoding: int
    values: np.ndarray                    # uint64 raw bits per record
    nulls: Optional[np.ndarray] = None    # bool per record -> optional column (dlevel_max = 1)
    bitpack_max: int = 0xFFFFFFFF         # column_writer_uint.cc:57-63: data pages always use the default max
    strings: Optional[Sequence[bytes]] = None   # STRING_PLAIN columns: one byte string per record (ignored where NULL)


def _paginate(payload: np.ndarray, bitpacked: bool, maxv: int) -> List[bytes]:
    pages = []
    if bitpacked:
        b = bits(maxv)
        if b == 0:
            return pages
        psz = 16 * b * 1024
        raw = payload.tobytes()
        for k in range(0, max(len(raw), 1), psz):
            chunk = raw[k:k + psz]
            chunk = chunk + b"\0" * (psz - len(chunk))
            if k == 0:
                chunk = struct.pack("<I", maxv) + chunk
            pages.append(chunk)
        if not raw:
            pages = []
    else:
        raw = payload.tobytes()
        for k in range(0, len(raw), PAGE_SIZE):
            chunk = raw[k:k + PAGE_SIZE]
            pages.append(chunk + b"\0" * (PAGE_SIZE - len(chunk)))
    return pages


def encode_data(encoding: int, vals: np.ndarray, maxv: int) -> np.ndarray:
    if encoding == P.ENC_UINT64_LEB128:
        return leb128_encode(vals)
    if encoding in (P.ENC_UINT64_PLAIN, P.ENC_FLOAT_IEEE754):
        return np.asarray(vals, dtype="<u8").view(np.uint8)
    if encoding == P.ENC_UINT32_PLAIN:
        return np.asarray(vals, dtype=np.uint64).astype("<u4").view(np.uint8)
    if encoding in (P.ENC_UINT32_BITPACKED, P.ENC_BOOLEAN_BITPACKED):
        return bitpack_vertical(vals, bits(maxv))
    raise OracleError("unsupported encoding %d" % encoding)


def write_cstable(path: str, num_rows: int, columns: Sequence[WriteColumn], interleave: bool = True) -> Dict[str, dict]:
    """Write a v0.2.0 file the reference reader accepts.  Returns per-column payload byte counts
    (the 'algorithmic bytes' sidecar of SURVEY §8(d))."""
    hdr = bytearray()
    hdr += MAGIC + struct.pack("<HQ", 2, 0)
    assert len(hdr) == META_POS
    hdr += b"\0" * (2 * META_SIZE) + b"\0" * 128
    hdr += write_varuint(len(columns))
    streams = []   # (kind, column_id, [pages])
    sidecar = {}
    for cid, c in enumerate(columns, start=1):
        optional = c.nulls is not None
        name = c.name.encode()
        hdr += write_varuint(c.logical_type) + write_varuint(c.encoding) + write_varuint(cid)
        hdr += write_varuint(len(name)) + name + write_varuint(0) + write_varuint(1 if optional else 0)
        is_string = c.encoding == P.ENC_STRING_PLAIN
        vals = np.zeros(num_rows, dtype=np.uint64) if is_string else np.asarray(c.values, dtype=np.uint64)
        assert len(vals) == num_rows
        level_bytes = 0
        if optional:
            nulls = np.asarray(c.nulls, dtype=bool)
            dl = (~nulls).astype(np.uint64)
            dpay = bitpack_vertical(dl, 1)
            level_bytes = ((num_rows + 127) // 128) * 16
            streams.append((P.STREAM_DLEVEL, cid, _paginate(dpay, True, 1)))
            vals = vals[~nulls]
        bitpacked = c.encoding in (P.ENC_UINT32_BITPACKED, P.ENC_BOOLEAN_BITPACKED)
        maxv = 1 if c.encoding == P.ENC_BOOLEAN_BITPACKED else c.bitpack_max
        if c.encoding == P.ENC_BOOLEAN_BITPACKED:
            vals = (vals > 0).astype(np.uint64)
        if is_string:
            # columns/page_writer_lenencstring.cc:36-48: varuint length + bytes, values run across page boundaries
            keep = np.ones(num_rows, dtype=bool) if not optional else ~np.asarray(c.nulls, dtype=bool)
            blob = bytearray()
            for i in range(num_rows):
                if keep[i]:
                    blob += write_varuint(len(c.strings[i])) + c.strings[i]
            pay = np.frombuffer(bytes(blob), dtype=np.uint8)
        else:
            pay = encode_data(c.encoding, vals, maxv)
        streams.append((P.STREAM_DATA, cid, _paginate(pay, bitpacked, maxv)))
        sidecar[c.name] = {"data_bytes": int(len(pay)), "level_bytes": int(level_bytes), "num_values": int(len(vals))}
    pad = (-len(hdr)) % SECTOR
    hdr += b"\0" * pad
    index = []
    body = bytearray()
    off = len(hdr)
    if interleave:
        # pages of different streams interleave as they fill, like the reference writer produces
        k = 0
        remaining = True
        while remaining:
            remaining = False
            for kind, cid, pages in streams:
                if k < len(pages):
                    index.append((kind, cid, off, len(pages[k])))
                    body += pages[k]
                    off += len(pages[k])
                    remaining = True
            k += 1
    else:
        for kind, cid, pages in streams:
            for pg in pages:
                index.append((kind, cid, off, len(pg)))
                body += pg
                off += len(pg)
    idx = bytearray(write_varuint(len(index)))
    for t, c, o, s in index:
        idx += write_varuint(t) + write_varuint(c) + write_varuint(o) + write_varuint(s)
    index_offset = off
    mb = struct.pack("<QQQI", 1, num_rows, index_offset, len(idx))
    mb += hashlib.sha1(mb).digest()
    # first commit has transaction id 1 -> metablock slot 1 (cstable_file.cc:172-174)
    hdr[META_POS + META_SIZE: META_POS + 2 * META_SIZE] = mb
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(body)
        f.write(idx)
    return sidecar


# ------------------------------------------------------------------------------------------------
# expression evaluation (sql/runtime/vm.cc:107-157 over sql/expressions/*.cc)
# ------------------------------------------------------------------------------------------------

@dataclass
class Vec:
    """A batch of values of one SType: raw values + STag bytes (sql/svalue.cc:533-549 packed elements)."""
    type: int
    values: np.ndarray     # uint64 | int64 | float64 | bool
    tags: np.ndarray       # uint8

    def bits64(self) -> np.ndarray:
        if self.type == P.STRING:
            # group identity of a string key is its bytes (groupby.cc:112-135 hashes them); any injective id will do
            _, inv = np.unique(self.values, return_inverse=True)
            return inv.astype(np.uint64)
        if self.type == P.FLOAT64:
            return self.values.view(np.uint64)
        if self.type == P.INT64:
            return self.values.view(np.uint64)
        if self.type == P.BOOL:
            return self.values.astype(np.uint64)
        return self.values.astype(np.uint64)


_NP_OF = {P.UINT64: np.uint64, P.TIMESTAMP64: np.uint64, P.INT64: np.int64, P.FLOAT64: np.float64, P.BOOL: np.bool_}

# datetime.cc:58-84
_US = {"ms": 1000, "msec": 1000, "millisecond": 1000, "milliseconds": 1000,
       "s": 10**6, "sec": 10**6, "second": 10**6, "seconds": 10**6,
       "min": 60 * 10**6, "minute": 60 * 10**6, "minutes": 60 * 10**6,
       "h": 3600 * 10**6, "hour": 3600 * 10**6, "hours": 3600 * 10**6,
       "d": 86400 * 10**6, "day": 86400 * 10**6, "days": 86400 * 10**6,
       "w": 7 * 86400 * 10**6, "week": 7 * 86400 * 10**6, "weeks": 7 * 86400 * 10**6,
       "month": 30 * 86400 * 10**6, "months": 30 * 86400 * 10**6,
       "y": 365 * 86400 * 10**6, "year": 365 * 86400 * 10**6, "years": 365 * 86400 * 10**6}


def date_trunc_window(window: str) -> int:
    """datetime.cc:115-137: std::stoull prefix as multiplier (default 1), rest is the unit name."""
    i = 0
    s = window.lstrip()
    while i < len(s) and s[i].isdigit():
        i += 1
    mult = int(s[:i]) if i else 1
    unit = s[i:] if i else window
    if unit not in _US:
        raise OracleError("unknown time window %s" % window)
    return _US[unit] * mult


def _trunc_div(a: np.ndarray, b: np.ndarray):
    """C integer division / remainder (truncation toward zero) for int64."""
    q = a // b
    r = a - q * b
    fix = (r != 0) & ((a < 0) != (b < 0))
    q = q + fix.astype(np.int64)
    r = a - q * b
    return q, r


def _const(type_: int, imm, n: int) -> Vec:
    dt = _NP_OF[type_]
    return Vec(type_, np.full(n, imm, dtype=dt), np.zeros(n, dtype=np.uint8))


def eval_expr(e: P.Expr, inputs: Sequence[Vec], n: int, active: Optional[np.ndarray] = None,
              agg_value: Optional[Tuple[P.Call, Vec]] = None) -> Vec:
    """Evaluate an expression tree for n rows.  `active` marks the rows whose result is observable:
    integer div/mod by zero raises only there (X_CJUMP makes `if` lazy, compiler.cc:174-209)."""
    if active is None:
        active = np.ones(n, dtype=bool)
    if agg_value is not None and e is agg_value[0]:
        return agg_value[1]
    if isinstance(e, P.Col):
        v = inputs[e.index]
        if v.type != e.type and not ({v.type, e.type} <= {P.UINT64, P.TIMESTAMP64}):
            raise OracleError("column type mismatch")
        return Vec(e.type, v.values, v.tags)                      # X_INPUT keeps the tag (vm.cc:138-142)
    if isinstance(e, P.Lit):
        if e.type == P.STRING:
            val = e.value.encode() if isinstance(e.value, str) else bytes(e.value)
            arr = np.empty(n, dtype=object)
            arr[:] = [val] * n
            return Vec(P.STRING, arr, np.zeros(n, dtype=np.uint8))
        if e.type == P.NIL:
            return Vec(P.NIL, np.zeros(n, dtype=np.uint64), np.ones(n, dtype=np.uint8))
        return _const(e.type, e.value, n)
    if isinstance(e, P.If):
        c = eval_expr(e.cond, inputs, n, active, agg_value)
        cb = c.values.astype(bool)
        t = eval_expr(e.then, inputs, n, active & cb, agg_value)
        f = eval_expr(e.otherwise, inputs, n, active & ~cb, agg_value)
        return Vec(t.type, np.where(cb, t.values, f.values), np.where(cb, t.tags, f.tags).astype(np.uint8))
    assert isinstance(e, P.Call)
    name, sig = e.symbol.split("#", 1)
    argt = [P.TYPE_BY_NAME[t] for t in sig.split("/", 1)[1].split(";") if t]
    if P.is_aggregate_symbol(e.symbol):
        raise OracleError("aggregate call in a pure context: %s" % e.symbol)
    if name == "date_trunc":
        w = e.args[0]
        if not isinstance(w, P.Lit):
            raise OracleError("date_trunc window must be a literal")
        ts = eval_expr(e.args[1], inputs, n, active, agg_value)
        t = np.uint64(date_trunc_window(w.value))
        return Vec(P.TIMESTAMP64, (ts.values.astype(np.uint64) // t) * t, np.zeros(n, dtype=np.uint8))
    a = [eval_expr(x, inputs, n, active, agg_value) for x in e.args]
    zt = np.zeros(n, dtype=np.uint8)                               # pure functions push tag 0 (svalue.cc:950-958)
    T = argt[0] if argt else P.NIL
    dt = _NP_OF.get(T, np.uint64)
    x = a[0].values.astype(dt, copy=False) if a and T != P.STRING else None
    y = a[1].values.astype(dt, copy=False) if len(a) > 1 and T != P.STRING else None
    with np.errstate(all="ignore"):
        if name == "logical_and":
            return Vec(P.BOOL, x.astype(bool) & y.astype(bool), zt)
        if name == "logical_or":
            return Vec(P.BOOL, x.astype(bool) | y.astype(bool), zt)
        if name == "neg":
            return Vec(P.BOOL, ~x.astype(bool), zt)
        if T == P.STRING and name in ("eq", "neq"):
            # boolean.cc:235-257, 355-377: length + memcmp on the popped strings; the tag is dropped, so a NULL compares as ""
            import operator
            op = {"eq": operator.eq, "neq": operator.ne}[name]
            return Vec(P.BOOL, np.fromiter((op(p, q) for p, q in zip(a[0].values, a[1].values)), dtype=bool, count=n), zt)
        if T == P.STRING and name in ("lt", "lte", "gt", "gte"):
            # boolean.cc:439-710: strncmp over the shorter length - which stops at an embedded NUL byte - then the lengths
            def cmp3(p, q):
                for x, y in zip(p, q):
                    if x != y:
                        return -1 if x < y else 1
                    if x == 0:
                        break
                return 0
            rel = {"lt": lambda c, lp, lq: c < 0 or (c == 0 and lp < lq), "lte": lambda c, lp, lq: c < 0 or (c == 0 and lp <= lq),
                   "gt": lambda c, lp, lq: c > 0 or (c == 0 and lp > lq), "gte": lambda c, lp, lq: c > 0 or (c == 0 and lp >= lq)}[name]
            return Vec(P.BOOL, np.fromiter((rel(cmp3(p, q), len(p), len(q)) for p, q in zip(a[0].values, a[1].values)), dtype=bool, count=n), zt)
        if name in ("startswith", "endswith"):
            # expressions/string.cc:52-74 (StringUtil::beginsWith / endsWith): (string, affix)
            f = bytes.startswith if name == "startswith" else bytes.endswith
            return Vec(P.BOOL, np.fromiter((f(p, q) for p, q in zip(a[0].values, a[1].values)), dtype=bool, count=n), zt)
        if name in ("eq", "neq", "lt", "lte", "gt", "gte"):
            r = {"eq": np.equal, "neq": np.not_equal, "lt": np.less, "lte": np.less_equal,
                 "gt": np.greater, "gte": np.greater_equal}[name](x, y)
            return Vec(P.BOOL, r, zt)
        if name == "cmp":
            return Vec(P.INT64, (x > y).astype(np.int64) - (x < y).astype(np.int64), zt)
        if name in ("add", "sub", "mul"):
            r = {"add": np.add, "sub": np.subtract, "mul": np.multiply}[name](x, y)
            return Vec(T, r, zt)
        if name in ("div", "mod"):
            if T == P.FLOAT64:
                r = x / y if name == "div" else np.fmod(x, y)
                return Vec(T, r, zt)
            zero = (y == 0) & active
            if zero.any():
                raise OracleError("division by zero" if name == "div" else "modulo by zero")
            ysafe = np.where(y == 0, np.array(1, dtype=dt), y)
            if T == P.UINT64:
                r = x // ysafe if name == "div" else x % ysafe
            else:
                q, rem = _trunc_div(x, ysafe)
                r = q if name == "div" else rem
            return Vec(T, r, zt)
        if name == "pow":
            r = np.power(x.astype(np.float64), y.astype(np.float64))
            return Vec(T, r.astype(dt), zt)
        if name == "to_nil":
            return Vec(P.NIL, np.zeros(n, dtype=np.uint64), zt)
        if name == "to_int64":
            src = a[0].values
            if a[0].type == P.FLOAT64:
                return Vec(P.INT64, np.trunc(src).astype(np.int64), zt)
            if a[0].type == P.BOOL:
                return Vec(P.INT64, src.astype(np.int64), zt)
            return Vec(P.INT64, src.astype(np.uint64).view(np.int64), zt)
        if name == "to_timestamp64":
            src = a[0].values
            if a[0].type == P.FLOAT64:
                return Vec(P.TIMESTAMP64, np.trunc(src).astype(np.uint64), zt)
            return Vec(P.TIMESTAMP64, src.astype(np.int64).view(np.uint64), zt)
        if name == "from_timestamp":
            src = a[0].values
            if a[0].type == P.FLOAT64:
                return Vec(P.TIMESTAMP64, np.trunc(src * 1e6).astype(np.uint64), zt)
            return Vec(P.TIMESTAMP64, (src.astype(np.int64) * np.int64(10**6)).view(np.uint64), zt)
    raise OracleError("symbol not found: %s" % e.symbol)


# ------------------------------------------------------------------------------------------------
# FastCSTableScan + GroupByExpression
# ------------------------------------------------------------------------------------------------

@dataclass
class Result:
    types: List[int]
    columns: List[Vec]
    num_rows: int
    rows_scanned: int = 0
    rows_passed: int = 0

    def rows(self) -> List[tuple]:
        """Rows as python tuples (None for NULL) for order-insensitive comparison."""
        out = []
        cols = []
        for v in self.columns:
            vals = v.values.tolist()
            tags = v.tags.tolist()
            cols.append([None if (t & 1) else x for x, t in zip(vals, tags)])
        for i in range(self.num_rows):
            out.append(tuple(c[i] for c in cols))
        return out

    def packed(self, idx: int) -> bytes:
        return pack_svector(self.columns[idx])


def order_by(res: Result, specs: Sequence[Tuple[int, bool]]) -> Result:
    """OrderByExpression::execute (sql/statements/select/orderby.cc:58-160): sort the rows with one typed `cmp` per
    sort spec (values only: the pure cmp functions drop NULL tags, SURVEY H7, so a NULL compares as its value bits 0).
    specs = [(result column, descending)], most significant first.  Rows with equal keys keep their order here; the
    reference's std::sort leaves it unspecified."""
    if not specs:
        raise OracleError("can't execute ORDER BY: no sort specs")
    idx = list(range(res.num_rows))
    for col, desc in reversed(list(specs)):
        v = res.columns[col]
        vals = v.values.tolist()
        if v.type == P.FLOAT64:
            vals = [0.0 if x == 0.0 else x for x in vals]      # -0.0 == +0.0
        idx.sort(key=lambda i: vals[i], reverse=bool(desc))   # python's sort is stable, also with reverse=True
    take = np.asarray(idx, dtype=np.int64)
    cols = [Vec(v.type, v.values[take], v.tags[take]) for v in res.columns]
    return Result(res.types, cols, res.num_rows, res.rows_scanned, res.rows_passed)


def limit(res: Result, count: int, offset: int = 0) -> Result:
    """LimitExpression::nextBatch (sql/statements/select/limit.cc:43-112): rows [offset, offset + count)."""
    lo = min(offset, res.num_rows)
    hi = min(lo + count, res.num_rows)
    cols = [Vec(v.type, v.values[lo:hi], v.tags[lo:hi]) for v in res.columns]
    return Result(res.types, cols, hi - lo, res.rows_scanned, res.rows_passed)


def pack_svector(v: Vec) -> bytes:
    """Packed SVector bytes (sql/svalue.cc:533-549): numeric [8 B value][1 B tag], BOOL [1 B][1 B tag]."""
    n = len(v.tags)
    if v.type == P.BOOL:
        out = np.zeros((n, 2), dtype=np.uint8)
        out[:, 0] = v.values.astype(np.uint8)
        out[:, 1] = v.tags
        return out.tobytes()
    out = np.zeros((n, 9), dtype=np.uint8)
    out[:, :8] = np.ascontiguousarray(v.bits64()).view(np.uint8).reshape(n, 8)
    out[:, 8] = v.tags
    return out.tobytes()


def unpack_svector(buf: bytes, type_: int, n: int) -> Vec:
    raw = np.frombuffer(buf, dtype=np.uint8)
    if type_ == P.BOOL:
        raw = raw[: 2 * n].reshape(n, 2)
        return Vec(P.BOOL, raw[:, 0].astype(bool), raw[:, 1].copy())
    raw = raw[: 9 * n].reshape(n, 9)
    bits_ = np.ascontiguousarray(raw[:, :8]).view("<u8").reshape(n)
    dt = _NP_OF[type_]
    return Vec(type_, bits_.view(dt) if dt != np.uint64 else bits_.copy(), raw[:, 8].copy())


def load_inputs(tables: Sequence[CSTableFile], names: Sequence[str]) -> Tuple[List[Vec], int]:
    """FastCSTableScan::fetchColumn* over every partition, concatenated (sql/CSTableScan.cc:860-968)."""
    vecs = []
    n_total = sum(t.num_rows for t in tables)
    for name in names:
        parts_v, parts_t, st = [], [], None
        for t in tables:
            if name not in t.columns:
                raise OracleError("column not found: %s" % name)
            if t.columns[name].logical_type == P.COL_STRING:
                sv = decode_string_column(t, name)           # fetchColumnString: NULL -> length 0 + STAG_NULL
                arr = np.empty(len(sv), dtype=object)
                arr[:] = [b"" if x is None else x for x in sv]
                st = P.STRING
                parts_v.append(arr)
                parts_t.append(np.array([1 if x is None else 0 for x in sv], dtype=np.uint8))
                continue
            d = decode_column(t, name)
            st = d.sql_type
            parts_v.append(d.values)
            parts_t.append(np.where(d.present, 0, 1).astype(np.uint8))
        raw = np.concatenate(parts_v) if parts_v else np.zeros(0, dtype=np.uint64)
        tags = np.concatenate(parts_t) if parts_t else np.zeros(0, dtype=np.uint8)
        if st == P.STRING:
            vecs.append(Vec(st, raw, tags))
            continue
        if st == P.FLOAT64:
            vals = raw.view(np.float64)
        elif st == P.BOOL:
            vals = raw.astype(bool)
        else:
            vals = raw
        vecs.append(Vec(st, vals, tags))
    return vecs, n_total


def _take(v: Vec, idx) -> Vec:
    return Vec(v.type, v.values[idx], v.tags[idx])


def run_query(tables: Sequence[CSTableFile], plan: P.QueryPlan, row_filter: Optional[np.ndarray] = None) -> Result:
    inputs, n = load_inputs(tables, plan.input_columns)
    return run_query_on(inputs, n, plan, row_filter)


def run_query_on(inputs: Sequence[Vec], n: int, plan: P.QueryPlan, row_filter: Optional[np.ndarray] = None) -> Result:
    """FastCSTableScan::nextBatch (sql/CSTableScan.cc:757-858): WHERE over every row, AND with the external
    filter, then (aggregate plans) GroupByExpression::execute (statements/select/groupby.cc:69-185)."""
    if plan.where is not None:
        w = eval_expr(plan.where, inputs, n)
        if w.type != P.BOOL:
            raise OracleError("WHERE must be BOOL")
        keep = w.values.astype(bool)                 # popBool drops the tag (H7)
    else:
        keep = np.ones(n, dtype=bool)
    if row_filter is not None:
        keep = keep & np.asarray(row_filter, dtype=bool)
    sel = np.flatnonzero(keep)
    m = len(sel)
    finputs = [_take(v, sel) for v in inputs]
    types = [s.type for s in plan.select]

    if not plan.is_groupby:
        cols = [eval_expr(s, finputs, m) for s in plan.select]
        return Result(types, cols, m, n, m)

    # ---- group identity: raw bytes of the evaluated key tuple incl. tags (groupby.cc:112-135) ----
    keys = [eval_expr(g, finputs, m) for g in plan.group]
    if m == 0:
        return Result(types, [Vec(t, np.zeros(0, dtype=_NP_OF.get(t, np.uint64)), np.zeros(0, dtype=np.uint8)) for t in types], 0, n, 0)
    if keys:
        sort_cols = []
        for k in keys:
            sort_cols.append(k.tags)
            sort_cols.append(k.bits64())
        order = np.lexsort(sort_cols[::-1])          # stable: first row of a group stays first
        boundary = np.zeros(m, dtype=bool)
        boundary[0] = True
        for c in sort_cols:
            cs = c[order]
            boundary[1:] |= cs[1:] != cs[:-1]
        starts = np.flatnonzero(boundary)
    else:
        order = np.arange(m)
        starts = np.array([0])
    ng = len(starts)
    sinputs = [_take(v, order) for v in finputs]
    first = starts

    out_cols = []
    for s in plan.select:
        agg = P.find_aggregate(s)
        if agg is None:
            # non-aggregate item: value on the group's first row (groupby.cc:161-172)
            full = eval_expr(s, sinputs, m)
            out_cols.append(_take(full, first))
            continue
        aname = agg.name
        argv = eval_expr(agg.args[0], sinputs, m) if agg.args else None
        res = _aggregate(aname, agg.type, argv, starts, m)
        if s is agg:
            out_cols.append(res)
        else:
            out_cols.append(eval_expr(s, [], ng, None, (agg, res)))
    return Result(types, out_cols, ng, n, m)


def _varuint(v: int) -> bytes:
    """OutputStream::appendVarUInt (util/io/outputstream.cc): unsigned LEB128."""
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def run_partial_query(tables: Sequence[CSTableFile], plan: P.QueryPlan) -> List[Tuple[bytes, bytes]]:
    """PartialGroupByExpression (sql/statements/select/groupby.cc:223-472), the shard side of a cluster GROUP BY: one row per
    group, (key, data).  key = SHA-1 of the group expressions' packed stack bytes, last expression first (groupby.cc:112-135:
    [8 B value][1 B tag], BOOL [1 B][1 B]).  data = the select items in order: an aggregate item as its function's saved
    state (count / sum_uint64 / sum_int64: varuint, aggregate.cc:52-58, 200-206; min / max {value, seen}, mean {double sum,
    n}, sum<float64> raw: oracle/ref_tools/ext_aggregates.cc), any other item as SValue::encode (svalue.cc:306-309)."""
    import hashlib
    import struct
    if not plan.is_groupby:
        raise OracleError("partial aggregation needs an aggregate plan")
    inputs, n = load_inputs(tables, plan.input_columns)
    if plan.where is not None:
        keep = eval_expr(plan.where, inputs, n).values.astype(bool)
    else:
        keep = np.ones(n, dtype=bool)
    sel = np.flatnonzero(keep)
    m = len(sel)
    if m == 0:
        return []
    finputs = [_take(v, sel) for v in inputs]
    keys = [eval_expr(g, finputs, m) for g in plan.group]
    if keys:
        sort_cols = []
        for k in keys:
            sort_cols.append(k.tags)
            sort_cols.append(k.bits64())
        order = np.lexsort(sort_cols[::-1])
        boundary = np.zeros(m, dtype=bool)
        boundary[0] = True
        for c in sort_cols:
            cs = c[order]
            boundary[1:] |= cs[1:] != cs[:-1]
        starts = np.flatnonzero(boundary)
    else:
        order = np.arange(m)
        starts = np.array([0])
    ng = len(starts)
    sinputs = [_take(v, order) for v in finputs]
    gkeys = [_take(_take(k, order), starts) for k in keys]

    def packed(v: Vec, i: int) -> bytes:
        if v.type == P.BOOL:
            return bytes([1 if v.values[i] else 0, int(v.tags[i])])
        if v.type == P.STRING:                   # [u32 length][bytes][tag] (svalue.cc:1139-1177); a NULL has length 0
            s = b"" if int(v.tags[i]) & 1 else bytes(v.values[i])
            return struct.pack("<I", len(s)) + s + bytes([int(v.tags[i])])
        return struct.pack("<Q", int(v.bits64()[i])) + bytes([int(v.tags[i])])

    items = []                                   # per select item: list of ng byte strings
    counts = np.diff(np.append(starts, m)).astype(np.uint64)
    for s in plan.select:
        agg = P.find_aggregate(s)
        if agg is None:
            col = _take(eval_expr(s, sinputs, m), starts)
            enc = []
            for i in range(ng):
                b = packed(col, i)
                # a packed value of exactly SValue::kInlineDataSize = 16 bytes (a string of 11 bytes) carries STAG_INLINE in
                # its tag byte: the packed tag and the SValue's own flag byte are the same byte (svalue.cc:346-365)
                if col.type == P.STRING and len(b) == 16:
                    b = b[:-1] + bytes([b[-1] | 128])
                enc.append(bytes([col.type]) + _varuint(len(b)) + b)
            items.append(enc)
            continue
        name = agg.name
        if name == "count":
            items.append([_varuint(int(c)) for c in counts])
            continue
        arg = eval_expr(agg.args[0], sinputs, m)
        if name == "count_distinct":
            # count_distinct_uint64_save (aggregate.cc:110-116): the size of the std::set, then its members in set order; a
            # NULL argument counts as its value 0 (popUInt64 drops the tag, H7)
            vals = arg.bits64()
            ends = np.append(starts[1:], m)
            enc = []
            for a, b in zip(starts, ends):
                members = np.unique(vals[a:b])
                enc.append(_varuint(len(members)) + b"".join(_varuint(int(v)) for v in members))
            items.append(enc)
            continue
        present = (arg.tags & 1) == 0
        npres = np.add.reduceat(present.astype(np.uint64), starts)
        res = _aggregate(name, agg.type, arg, starts, m)
        if name == "sum":
            if agg.type == P.FLOAT64:
                items.append([struct.pack("<d", float(x)) for x in res.values])
            else:
                items.append([_varuint(int(x)) for x in res.bits64()])
        elif name in ("min", "max"):
            bits = res.bits64()
            items.append([struct.pack("<QQ", int(bits[i]) if npres[i] else 0, 1 if npres[i] else 0) for i in range(ng)])
        elif name == "mean":
            with np.errstate(all="ignore"):
                sums = np.add.reduceat(np.where(present, arg.values.astype(np.float64), 0.0), starts)
            items.append([struct.pack("<dQ", float(sums[i]), int(npres[i])) for i in range(ng)])
        else:
            raise OracleError("unknown aggregate %s" % name)
    out = []
    for i in range(ng):
        kb = b"".join(packed(k, i) for k in reversed(gkeys))
        out.append((hashlib.sha1(kb).digest(), b"".join(it[i] for it in items)))
    return out


def _read_varuint_at(data: bytes, pos: int) -> Tuple[int, int]:
    v, sh = 0, 0
    while True:
        b = data[pos]
        pos += 1
        v |= (b & 0x7F) << sh
        sh += 7
        if not b & 0x80:
            return v, pos


def parse_partial_states(plan: P.QueryPlan, data: bytes) -> List[tuple]:
    """The saved states of one partial-aggregation row, one entry per select item, as VM::loadInstanceState / SValue::decode
    read them (groupby.cc:262-292, 590-600): ("value", type, packed bytes) | ("count", n) | ("sum", raw 64-bit or float) |
    ("minmax", raw 64-bit value, seen) | ("mean", float sum, n)."""
    out, pos = [], 0
    for s in plan.select:
        agg = P.find_aggregate(s)
        if agg is None:
            t = data[pos]
            n, pos = _read_varuint_at(data, pos + 1)
            out.append(("value", t, data[pos:pos + n]))
            pos += n
        elif agg.name == "count":
            v, pos = _read_varuint_at(data, pos)
            out.append(("count", v))
        elif agg.name == "sum" and agg.type != P.FLOAT64:
            v, pos = _read_varuint_at(data, pos)
            out.append(("sum", v))
        elif agg.name == "sum":
            out.append(("sum", struct.unpack_from("<d", data, pos)[0]))
            pos += 8
        elif agg.name in ("min", "max"):
            v, seen = struct.unpack_from("<QQ", data, pos)
            out.append(("minmax", v, seen))
            pos += 16
        elif agg.name == "mean":
            sm, n = struct.unpack_from("<dQ", data, pos)
            out.append(("mean", sm, n))
            pos += 16
        elif agg.name == "count_distinct":       # count_distinct_uint64_load (aggregate.cc:118-124)
            n, pos = _read_varuint_at(data, pos)
            members = set()
            for _ in range(n):
                v, pos = _read_varuint_at(data, pos)
                members.add(v)
            out.append(("distinct", members))
        else:
            raise OracleError("aggregate %s has no partial state format" % agg.name)
    if pos != len(data):
        raise OracleError("invalid partialaggr result encoding")
    return out


def merge_partial_rows(plan: P.QueryPlan, row_lists: Sequence[Sequence[Tuple[bytes, bytes]]]) -> List[tuple]:
    """GroupByMergeExpression (sql/statements/select/groupby.cc:553-615 execute, :617-660 nextBatch): the coordinator loads
    the shards' (group key, saved states) rows, merges the states of equal keys with the aggregates' merge functions
    (count / sum: += aggregate.cc:48-50, 196-198; min / max: over the seen ones, mean: sums and counts add,
    sum<float64>: += - oracle/ref_tools/ext_aggregates.cc) - a non-aggregate item is overwritten by every row that carries
    it (SValue::decode) - and evaluates every select item's `get` side per group.  Returns rows as python tuples (None = NULL)."""
    M64 = (1 << 64) - 1
    groups: Dict[bytes, list] = {}
    for rows in row_lists:
        for key, data in rows:
            st = [list(x) for x in parse_partial_states(plan, data)]
            cur = groups.get(key)
            if cur is None:
                groups[key] = st
                continue
            for i, (s, a, b) in enumerate(zip(plan.select, cur, st)):
                kind = a[0]
                if kind == "value":
                    cur[i] = b
                elif kind == "count":
                    a[1] = (a[1] + b[1]) & M64
                elif kind == "sum":
                    a[1] = (a[1] + b[1]) & M64 if isinstance(a[1], int) else a[1] + b[1]
                elif kind == "distinct":         # count_distinct_uint64_merge (aggregate.cc:102-108): the union
                    a[1] |= b[1]
                elif kind == "mean":
                    a[1] += b[1]
                    a[2] = (a[2] + b[2]) & M64
                elif kind == "minmax":
                    if not b[2]:
                        continue
                    agg = P.find_aggregate(s)
                    dt = {P.INT64: "<q", P.FLOAT64: "<d"}.get(agg.type, "<Q")
                    va = struct.unpack(dt, struct.pack("<Q", a[1]))[0]
                    vb = struct.unpack(dt, struct.pack("<Q", b[1]))[0]
                    if not a[2] or (vb > va if agg.name == "max" else vb < va):
                        a[1] = b[1]
                    a[2] = 1
    out = []
    zt = np.zeros(1, dtype=np.uint8)
    for key, st in groups.items():
        row = []
        for s, a in zip(plan.select, st):
            agg = P.find_aggregate(s)
            if agg is None:
                t, raw = a[1], a[2]
                if t == P.BOOL:
                    row.append(None if raw[1] & 1 else bool(raw[0]))
                else:
                    bits_ = struct.unpack_from("<Q", raw, 0)[0]
                    if raw[8] & 1:
                        row.append(None)
                    elif t == P.FLOAT64:
                        row.append(struct.unpack("<d", struct.pack("<Q", bits_))[0])
                    elif t == P.INT64:
                        row.append(struct.unpack("<q", struct.pack("<Q", bits_))[0])
                    else:
                        row.append(bits_)
                continue
            if a[0] == "count":
                res = Vec(P.UINT64, np.array([a[1]], dtype=np.uint64), zt)
            elif a[0] == "distinct":
                res = Vec(P.UINT64, np.array([len(a[1])], dtype=np.uint64), zt)
            elif a[0] == "sum":
                if agg.type == P.FLOAT64:
                    res = Vec(P.FLOAT64, np.array([a[1]], dtype=np.float64), zt)
                else:
                    res = Vec(agg.type, np.array([a[1]], dtype=np.uint64).view(_NP_OF[agg.type]), zt)
            elif a[0] == "minmax":
                res = Vec(agg.type, np.array([a[1]], dtype=np.uint64).view(_NP_OF[agg.type]), zt)     # MinMax::get: the value (0 if none seen)
            else:
                with np.errstate(all="ignore"):
                    res = Vec(P.FLOAT64, np.array([a[1]], dtype=np.float64) / np.float64(a[2]), zt)   # Mean::get: sum / n
            v = res if s is agg else eval_expr(s, [], 1, None, (agg, res))
            val = v.values.tolist()[0]
            row.append(None if int(v.tags[0]) & 1 else val)
        out.append(tuple(row))
    return out


def _aggregate(name: str, rtype: int, arg: Optional[Vec], starts: np.ndarray, m: int) -> Vec:
    ng = len(starts)
    zt = np.zeros(ng, dtype=np.uint8)
    counts = np.diff(np.append(starts, m)).astype(np.uint64)
    if name == "count":                                   # aggregate.cc:35-71: ++ for every row
        return Vec(P.UINT64, counts, zt)
    present = (arg.tags & 1) == 0
    with np.errstate(all="ignore"):
        if name == "sum":                                 # aggregate.cc:184-219: acc += v, NULL carries value 0
            if rtype == P.FLOAT64:
                return Vec(P.FLOAT64, np.add.reduceat(arg.values.astype(np.float64), starts), zt)
            return Vec(rtype, np.add.reduceat(arg.values, starts), zt)
        if name in ("min", "max"):                        # ext_aggregates.cc MinMax: NULLs skipped, 0 if none seen
            v = arg.values
            if name == "min":
                fill = np.inf if v.dtype == np.float64 else np.iinfo(v.dtype).max
                r = np.minimum.reduceat(np.where(present, v, np.array(fill, dtype=v.dtype)), starts)
            else:
                fill = -np.inf if v.dtype == np.float64 else np.iinfo(v.dtype).min
                r = np.maximum.reduceat(np.where(present, v, np.array(fill, dtype=v.dtype)), starts)
            seen = np.add.reduceat(present.astype(np.uint64), starts) > 0
            r = np.where(seen, r, np.zeros(1, dtype=v.dtype))
            return Vec(rtype, r.astype(v.dtype), zt)
        if name == "count_distinct":                      # aggregate.cc:80-137: std::set of the VALUES (a NULL is its value 0)
            ends = np.append(starts, m)
            r = np.array([len(np.unique(arg.values[ends[i]:ends[i + 1]])) for i in range(ng)], dtype=np.uint64)
            return Vec(P.UINT64, r, zt)
        if name == "mean":                                # ext_aggregates.cc Mean: sum(double)/n over non-NULL
            v = np.where(present, arg.values.astype(np.float64), 0.0)
            s = np.add.reduceat(v, starts)
            c = np.add.reduceat(present.astype(np.uint64), starts)
            return Vec(P.FLOAT64, s / c.astype(np.float64), zt)
    raise OracleError("unknown aggregate %s" % name)
