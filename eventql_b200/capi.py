"""ctypes binding of the C ABI in include/evqgpu.h (libevqgpu.so).

This is the binding a Python-side maintainer of the reference's drivers would add; the tests and
bench.py call the CUDA path through it.  There is no fallback: if the library is missing or has no
device, every entry point raises EvqError.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import plan as P

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libevqgpu.so")

STATUS_NAMES = {0: "OK", 1: "ERR_ARG", 2: "ERR_UNSUPPORTED", 3: "ERR_CUDA", 4: "ERR_RUNTIME", 5: "ERR_FORMAT", 6: "ERR_NOMEM"}


class EvqError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status
        self.message = message


class Insn(C.Structure):
    _fields_ = [("op", C.c_uint8), ("type", C.c_uint8), ("nargs", C.c_uint16), ("arg", C.c_uint32), ("imm", C.c_uint64)]


class ExprC(C.Structure):
    _fields_ = [("code", C.POINTER(Insn)), ("len", C.c_uint32), ("strings", C.c_char_p), ("strings_len", C.c_uint32)]


class QueryDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("num_input_columns", C.c_uint32),
                ("input_columns", C.POINTER(C.c_char_p)), ("where", ExprC), ("num_group", C.c_uint32),
                ("group", C.POINTER(ExprC)), ("num_select", C.c_uint32), ("select", C.POINTER(ExprC)),
                ("expected_groups", C.c_uint64)]


class ColumnInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("column_id", C.c_uint32), ("logical_type", C.c_uint32), ("encoding", C.c_uint32),
                ("rlevel_max", C.c_uint32), ("dlevel_max", C.c_uint32), ("sql_type", C.c_uint32), ("loaded", C.c_uint32),
                ("data_bytes", C.c_uint64), ("level_bytes", C.c_uint64), ("num_values", C.c_uint64),
                ("value_bits", C.c_uint32), ("leb_max_len", C.c_uint32), ("value_min", C.c_uint64), ("value_max", C.c_uint64)]


class SynthColumn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("logical_type", C.c_uint32), ("encoding", C.c_uint32), ("null_every", C.c_uint32),
                ("transform", C.c_uint32), ("seed", C.c_uint64), ("lo", C.c_uint64), ("span", C.c_uint64)]


class QueryStats(C.Structure):
    _fields_ = [("rows_scanned", C.c_uint64), ("rows_passed", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("num_groups", C.c_uint64), ("kernel_launches", C.c_uint32), ("strategy", C.c_uint32),
                ("jit_ms", C.c_float), ("scan_ms", C.c_float), ("scan_launches", C.c_uint32), ("jit_disk_hits", C.c_uint32)]


class SortSpec(C.Structure):
    _fields_ = [("column", C.c_uint32), ("descending", C.c_uint32)]


class LsmSegment(C.Structure):
    _fields_ = [("table", C.c_void_p), ("skiplist", C.c_void_p), ("flags", C.c_uint32), ("filtered", C.c_uint32),
                ("visible_rows", C.c_uint64)]


LSM_SKIP_COLUMN = 1
LSM_NO_FILTER = 2
LSM_HAS_UPDATES = 4
LSM_OLDEST = 8
LSM_AUTO = 16


class DebugColumn(C.Structure):
    _fields_ = [("sql_type", C.c_uint32), ("encoding", C.c_uint32), ("dlevel_max", C.c_uint32), ("value_bits", C.c_uint32),
                ("leb_max_len", C.c_uint32), ("reserved", C.c_uint32), ("value_min", C.c_uint64), ("value_max", C.c_uint64)]


# every symbol include/evqgpu.h declares (tests/test_abi.py checks header and library against this list)
SYMBOLS = [
    "evqgpu_ctx_create", "evqgpu_ctx_destroy", "evqgpu_last_error", "evqgpu_abi_version", "evqgpu_host_alloc",
    "evqgpu_host_free", "evqgpu_host_register", "evqgpu_host_unregister", "evqgpu_ctx_stream", "evqgpu_ctx_synchronize",
    "evqgpu_ctx_set_profiling", "evqgpu_ctx_kernel_launches",
    "evqgpu_table_open", "evqgpu_table_create", "evqgpu_table_add_column", "evqgpu_table_add_stream",
    "evqgpu_table_set_filter", "evqgpu_table_destroy", "evqgpu_table_num_rows", "evqgpu_table_num_columns", "evqgpu_table_column_info",
    "evqgpu_table_find_column", "evqgpu_table_load_columns", "evqgpu_table_read_stream", "evqgpu_table_decode_column",
    "evqgpu_table_write_file", "evqgpu_table_synthesize", "evqgpu_function_lookup", "evqgpu_function_symbol",
    "evqgpu_function_is_aggregate", "evqgpu_query_create", "evqgpu_query_destroy", "evqgpu_query_num_columns",
    "evqgpu_query_column_type", "evqgpu_query_prepare", "evqgpu_query_execute", "evqgpu_query_enqueue", "evqgpu_query_finish",
    "evqgpu_query_num_rows", "evqgpu_query_fetch", "evqgpu_query_order_by", "evqgpu_query_limit", "evqgpu_query_fetch_partial", "evqgpu_query_get_stats", "evqgpu_query_kernel_source",
    "evqgpu_comm_unique_id", "evqgpu_comm_init", "evqgpu_comm_destroy", "evqgpu_query_merge", "evqgpu_debug_generate",
    "evqgpu_table_decode_string_column", "evqgpu_table_get_filter", "evqgpu_lsm_build_filters", "evqgpu_query_fetch_strings",
    "evqgpu_partial_cache_encode", "evqgpu_partial_cache_filename", "evqgpu_query_store_cache",
    "evqgpu_partial_frames_encode", "evqgpu_partial_rows_split", "evqgpu_partial_cache_decode", "evqgpu_partial_frames_decode",
    "evqgpu_query_merge_rows", "evqgpu_query_merge_finish",
]

_lib = None


def lib() -> C.CDLL:
    """Load libevqgpu.so (built in-tree by eventql_b200.build). Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EvqError(3, "native library %s is missing - run `python -m eventql_b200.build`" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u32, u64, cp = C.c_void_p, C.c_uint32, C.c_uint64, C.c_char_p
    L.evqgpu_last_error.restype = cp
    L.evqgpu_ctx_create.argtypes = [C.c_int, u64, C.POINTER(vp)]
    L.evqgpu_ctx_destroy.argtypes = [vp]
    L.evqgpu_ctx_destroy.restype = None
    L.evqgpu_host_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.evqgpu_host_free.argtypes = [vp, vp]
    L.evqgpu_host_register.argtypes = [vp, vp, u64]
    L.evqgpu_host_unregister.argtypes = [vp, vp]
    L.evqgpu_ctx_stream.argtypes = [vp]
    L.evqgpu_ctx_stream.restype = vp
    L.evqgpu_ctx_synchronize.argtypes = [vp]
    L.evqgpu_ctx_set_profiling.argtypes = [vp, C.c_int]
    L.evqgpu_ctx_kernel_launches.argtypes = [vp]
    L.evqgpu_ctx_kernel_launches.restype = u64
    L.evqgpu_table_open.argtypes = [vp, vp, u64, C.POINTER(vp)]
    L.evqgpu_table_create.argtypes = [vp, u64, C.POINTER(vp)]
    L.evqgpu_table_add_column.argtypes = [vp, cp, u32, u32, u32, u32]
    L.evqgpu_table_add_stream.argtypes = [vp, cp, u32, vp, u64, u32, u32]
    L.evqgpu_table_set_filter.argtypes = [vp, vp, u64, u32]
    L.evqgpu_table_destroy.argtypes = [vp]
    L.evqgpu_table_destroy.restype = None
    L.evqgpu_table_num_rows.argtypes = [vp]
    L.evqgpu_table_num_rows.restype = u64
    L.evqgpu_table_num_columns.argtypes = [vp]
    L.evqgpu_table_num_columns.restype = u32
    L.evqgpu_table_column_info.argtypes = [vp, u32, C.POINTER(ColumnInfo)]
    L.evqgpu_table_find_column.argtypes = [vp, cp]
    L.evqgpu_table_load_columns.argtypes = [vp, C.POINTER(cp), u32]
    L.evqgpu_table_read_stream.argtypes = [vp, cp, u32, vp, u64, C.POINTER(u64), C.POINTER(u32)]
    L.evqgpu_table_decode_column.argtypes = [vp, cp, u64, u64, vp, u64]
    L.evqgpu_table_write_file.argtypes = [vp, cp]
    L.evqgpu_table_decode_string_column.argtypes = [vp, cp, u64, u64, vp, u64, C.POINTER(u64)]
    L.evqgpu_table_get_filter.argtypes = [vp, vp, u64, C.POINTER(C.c_int)]
    L.evqgpu_lsm_build_filters.argtypes = [vp, C.POINTER(LsmSegment), u32]
    L.evqgpu_table_synthesize.argtypes = [vp, u64, u64, C.POINTER(SynthColumn), u32, C.POINTER(vp)]
    L.evqgpu_function_lookup.argtypes = [cp]
    L.evqgpu_function_symbol.argtypes = [C.c_int]
    L.evqgpu_function_symbol.restype = cp
    L.evqgpu_function_is_aggregate.argtypes = [C.c_int]
    L.evqgpu_query_create.argtypes = [vp, C.POINTER(QueryDesc), C.POINTER(vp)]
    L.evqgpu_query_destroy.argtypes = [vp]
    L.evqgpu_query_destroy.restype = None
    L.evqgpu_query_num_columns.argtypes = [vp]
    L.evqgpu_query_num_columns.restype = u32
    L.evqgpu_query_column_type.argtypes = [vp, u32]
    L.evqgpu_query_column_type.restype = u32
    L.evqgpu_query_execute.argtypes = [vp, C.POINTER(vp), u32]
    L.evqgpu_query_enqueue.argtypes = [vp, C.POINTER(vp), u32]
    L.evqgpu_query_prepare.argtypes = [vp, C.POINTER(vp), u32]
    L.evqgpu_query_merge_rows.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(u64), u64]
    L.evqgpu_query_merge_finish.argtypes = [vp]
    L.evqgpu_query_finish.argtypes = [vp]
    L.evqgpu_query_num_rows.argtypes = [vp, C.POINTER(u64)]
    L.evqgpu_query_fetch.argtypes = [vp, u64, u64, C.POINTER(vp), C.POINTER(u64)]
    L.evqgpu_query_fetch_strings.argtypes = [vp, u32, u64, u64, vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.evqgpu_partial_cache_encode.argtypes = [vp, vp, C.POINTER(u64), u64, vp, u64, C.POINTER(u64)]
    L.evqgpu_partial_frames_encode.argtypes = [vp, vp, C.POINTER(u64), u64, u64, vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.evqgpu_partial_rows_split.argtypes = [C.POINTER(QueryDesc), vp, u64, C.POINTER(u64), u64, C.POINTER(u64)]
    L.evqgpu_partial_cache_decode.argtypes = [C.POINTER(QueryDesc), vp, u64, C.POINTER(u64), u64, C.POINTER(u64)]
    L.evqgpu_partial_frames_decode.argtypes = [C.POINTER(QueryDesc), vp, u64, C.POINTER(u64), u64, C.POINTER(u64), C.POINTER(u64),
                                               C.POINTER(C.c_int)]
    L.evqgpu_partial_cache_filename.argtypes = [vp, vp, C.c_char_p, u64]
    L.evqgpu_query_store_cache.argtypes = [vp, cp]
    L.evqgpu_query_order_by.argtypes = [vp, C.POINTER(SortSpec), u32]
    L.evqgpu_query_limit.argtypes = [vp, u64, u64]
    L.evqgpu_query_fetch_partial.argtypes = [vp, u64, u64, vp, vp, u64, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.evqgpu_query_get_stats.argtypes = [vp, C.POINTER(QueryStats)]
    L.evqgpu_query_kernel_source.argtypes = [vp]
    L.evqgpu_query_kernel_source.restype = cp
    L.evqgpu_comm_unique_id.argtypes = [vp]
    L.evqgpu_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.evqgpu_comm_destroy.argtypes = [vp]
    L.evqgpu_query_merge.argtypes = [vp]
    L.evqgpu_debug_generate.argtypes = [C.POINTER(QueryDesc), C.POINTER(DebugColumn), u32, u32, cp, u64, C.POINTER(u64), C.c_int,
                                        C.POINTER(u64)]
    _lib = L
    return L


def partial_cache_encode(rows) -> bytes:
    """[(20-byte group key, saved states)] -> the .qc query cache entry of PartialGroupByExpression (groupby.cc:411-432)."""
    n = len(rows)
    keys = b"".join(k for k, _ in rows)
    data = b"".join(d for _, d in rows)
    offs = (C.c_uint64 * (n + 1))()
    pos = 0
    for i, (_, d) in enumerate(rows):
        offs[i] = pos
        pos += len(d)
    offs[n] = pos
    need = C.c_uint64(0)
    kb = C.create_string_buffer(keys, max(1, len(keys)))
    db = C.create_string_buffer(data, max(1, len(data)))
    check(lib().evqgpu_partial_cache_encode(kb, db, offs, n, None, 0, C.byref(need)))
    out = C.create_string_buffer(max(1, need.value))
    check(lib().evqgpu_partial_cache_encode(kb, db, offs, n, out, need.value, C.byref(need)))
    return out.raw[: need.value]


def partial_frames_encode(rows, soft_max_body: int = 0) -> bytes:
    """[(20-byte group key, saved states)] -> the QUERY_PARTIALAGGR_RESULT frames of a shard's answer, back to back."""
    n = len(rows)
    keys = b"".join(k for k, _ in rows)
    data = b"".join(d for _, d in rows)
    offs = (C.c_uint64 * (n + 1))()
    pos = 0
    for i, (_, d) in enumerate(rows):
        offs[i] = pos
        pos += len(d)
    offs[n] = pos
    need, frames = C.c_uint64(0), C.c_uint64(0)
    kb = C.create_string_buffer(keys, max(1, len(keys)))
    db = C.create_string_buffer(data, max(1, len(data)))
    check(lib().evqgpu_partial_frames_encode(kb, db, offs, n, soft_max_body, None, 0, C.byref(need), C.byref(frames)))
    out = C.create_string_buffer(max(1, need.value))
    check(lib().evqgpu_partial_frames_encode(kb, db, offs, n, soft_max_body, out, need.value, C.byref(need), C.byref(frames)))
    return out.raw[: need.value]


def _split(fn, plan: P.QueryPlan, buf: bytes, pairs: bool = False, extra=()):
    pc = _PlanC(plan)
    raw = C.create_string_buffer(buf, max(1, len(buf)))
    n = C.c_uint64(0)
    check(fn(C.byref(pc.desc), raw, len(buf), None, 0, C.byref(n), *extra))
    offs = (C.c_uint64 * ((2 * n.value if pairs else n.value + 1) + 1))()
    check(fn(C.byref(pc.desc), raw, len(buf), offs, n.value, C.byref(n), *extra))
    if pairs:
        return [(buf[offs[2 * i]: offs[2 * i] + 20], buf[offs[2 * i] + 20: offs[2 * i + 1]]) for i in range(n.value)]
    return [(buf[offs[i]: offs[i] + 20], buf[offs[i] + 20: offs[i + 1]]) for i in range(n.value)]


def partial_rows_split(plan: P.QueryPlan, body: bytes):
    """`20-byte key | saved states` rows back to back -> [(key, states)], walked with the plan (GroupByMergeExpression's read loop)."""
    return _split(lib().evqgpu_partial_rows_split, plan, body)


def partial_cache_decode(plan: P.QueryPlan, entry: bytes):
    """A .qc query cache entry -> [(key, states)]."""
    return _split(lib().evqgpu_partial_cache_decode, plan, entry)


def partial_frames_decode(plan: P.QueryPlan, frames: bytes):
    """QUERY_PARTIALAGGR_RESULT frames back to back -> ([(key, states)], number of frames, end-of-request seen)."""
    nframes, eor = C.c_uint64(0), C.c_int(0)
    rows = _split(lib().evqgpu_partial_frames_decode, plan, frames, pairs=True, extra=(C.byref(nframes), C.byref(eor)))
    return rows, nframes.value, bool(eor.value)


def partial_cache_filename(input_cache_key: bytes, expression_fingerprint: bytes) -> str:
    out = C.create_string_buffer(44)
    check(lib().evqgpu_partial_cache_filename(input_cache_key, expression_fingerprint, out, 44))
    return out.value.decode()


def check(rc: int) -> None:
    if rc != 0:
        raise EvqError(rc, lib().evqgpu_last_error().decode(errors="replace"))


class _PlanC:
    """Keeps the ctypes buffers of one evqgpu_query_desc alive."""

    def __init__(self, plan: P.QueryPlan):
        L = lib()
        self._keep = []

        def fn_id(sym: str) -> int:
            return L.evqgpu_function_lookup(sym.encode())

        def expr(e: Optional[P.Expr]) -> ExprC:
            prog = P.flatten(e, fn_id)
            ec = ExprC()
            if prog.insns:
                arr = (Insn * len(prog.insns))()
                for i, (op, ty, nargs, arg, imm) in enumerate(prog.insns):
                    arr[i] = Insn(op, ty, nargs, arg, imm)
                self._keep.append(arr)
                ec.code = C.cast(arr, C.POINTER(Insn))
                ec.len = len(prog.insns)
                if prog.strings:
                    buf = C.create_string_buffer(prog.strings, len(prog.strings))
                    self._keep.append(buf)
                    ec.strings = C.cast(buf, C.c_char_p)
                    ec.strings_len = len(prog.strings)
            return ec

        d = QueryDesc()
        d.struct_size = C.sizeof(QueryDesc)
        d.flags = plan.flags
        names = (C.c_char_p * max(1, len(plan.input_columns)))(*[n.encode() for n in plan.input_columns])
        self._keep.append(names)
        d.num_input_columns = len(plan.input_columns)
        d.input_columns = C.cast(names, C.POINTER(C.c_char_p))
        d.where = expr(plan.where)
        groups = (ExprC * max(1, len(plan.group)))(*[expr(g) for g in plan.group])
        self._keep.append(groups)
        d.num_group = len(plan.group)
        d.group = C.cast(groups, C.POINTER(ExprC))
        sels = (ExprC * max(1, len(plan.select)))(*[expr(s) for s in plan.select])
        self._keep.append(sels)
        d.num_select = len(plan.select)
        d.select = C.cast(sels, C.POINTER(ExprC))
        d.expected_groups = plan.expected_groups
        self.desc = d


def debug_generate(plan: P.QueryPlan, columns: Sequence[tuple], tier: int = 1, dense_slots: int = 1, compile: bool = True):
    """Device-free: kernel text (and NVRTC compilation for sm_100a) of a plan over columns [(sql_type, encoding, dlevel_max[, value_bits, leb_max_len, 0, value_min, value_max])]."""
    L = lib()
    pc = _PlanC(plan)
    cols = (DebugColumn * max(1, len(columns)))(*[DebugColumn(*(tuple(c) + (0,) * (8 - len(c)))) for c in columns])
    n = C.c_uint64(0)
    cub = C.c_uint64(0)
    check(L.evqgpu_debug_generate(C.byref(pc.desc), cols, tier, dense_slots, None, 0, C.byref(n), 0, None))
    buf = C.create_string_buffer(n.value + 1)
    check(L.evqgpu_debug_generate(C.byref(pc.desc), cols, tier, dense_slots, buf, n.value + 1, C.byref(n), 1 if compile else 0, C.byref(cub)))
    return buf.value.decode(), cub.value


class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().evqgpu_ctx_create(device, 0, C.byref(self._h)))
        self.device = device
        self.rank, self.nranks = 0, 1

    def close(self):
        if self._h:
            lib().evqgpu_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self) -> int:
        return lib().evqgpu_ctx_stream(self._h) or 0

    def synchronize(self):
        check(lib().evqgpu_ctx_synchronize(self._h))

    def set_profiling(self, on: bool):
        check(lib().evqgpu_ctx_set_profiling(self._h, 1 if on else 0))

    @property
    def kernel_launches(self) -> int:
        return lib().evqgpu_ctx_kernel_launches(self._h)

    # ---- multi-GPU (one process per GPU; the launcher distributes the id, e.g. with torch.distributed)
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().evqgpu_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().evqgpu_comm_init(self._h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    # ---- tables
    def open_table(self, data) -> "Table":
        """data: bytes / numpy uint8 array holding a cstable file image (kept alive by the Table)."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        h = C.c_void_p()
        check(lib().evqgpu_table_open(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes, C.byref(h)))
        return Table(self, h, keep=arr)

    def open_table_file(self, path: str) -> "Table":
        return self.open_table(np.fromfile(path, dtype=np.uint8))

    def synthesize(self, num_rows: int, columns: Sequence[dict], row_offset: int = 0) -> "Table":
        arr = (SynthColumn * len(columns))()
        keep = []
        for i, c in enumerate(columns):
            name = c["name"].encode()
            keep.append(name)
            arr[i] = SynthColumn(name, c.get("logical_type", P.COL_UNSIGNED_INT), c.get("encoding", P.ENC_UINT64_LEB128),
                                 c.get("null_every", 0), c.get("transform", 0), c.get("seed", 0), c.get("lo", 0), c.get("span", 1))
        h = C.c_void_p()
        check(lib().evqgpu_table_synthesize(self._h, num_rows, row_offset, arr, len(columns), C.byref(h)))
        return Table(self, h)

    def create_table(self, num_rows: int) -> "Table":
        h = C.c_void_p()
        check(lib().evqgpu_table_create(self._h, num_rows, C.byref(h)))
        return Table(self, h)

    def query(self, plan: P.QueryPlan) -> "Query":
        return Query(self, plan)

    def lsm_build_filters(self, segments) -> List[int]:
        """PartitionCursor's visibility filters (server/sql/partition_cursor.cc:157-194) for the segments of one
        partition, in the cursor's order.  segments: (Table, skiplist bools or None, use_skip_column, needs_filter[,
        has_updates, oldest]); needs_filter None = the cursor's own rule (LSM_AUTO with has_updates / oldest).
        Installs each table's row filter; returns the visible rows per segment (self.lsm_filtered: which got a filter)."""
        arr = (LsmSegment * max(1, len(segments)))()
        keep = []
        for i, seg in enumerate(segments):
            tbl, skiplist, use_skip_column, needs_filter = seg[:4]
            has_updates, oldest = (seg[4], seg[5]) if len(seg) > 4 else (False, False)
            arr[i].table = tbl._h
            fl = LSM_SKIP_COLUMN if use_skip_column else 0
            if needs_filter is None:
                fl |= LSM_AUTO | (LSM_HAS_UPDATES if has_updates else 0) | (LSM_OLDEST if oldest else 0)
            elif not needs_filter:
                fl |= LSM_NO_FILTER
            arr[i].flags = fl
            if skiplist is not None:
                bits = np.packbits(np.asarray(skiplist, dtype=bool), bitorder="little")
                if bits.size == 0:
                    bits = np.zeros(1, dtype=np.uint8)
                keep.append(bits)
                arr[i].skiplist = bits.ctypes.data_as(C.c_void_p)
        check(lib().evqgpu_lsm_build_filters(self._h, arr, len(segments)))
        self.lsm_filtered = [bool(arr[i].filtered) for i in range(len(segments))]
        return [int(arr[i].visible_rows) for i in range(len(segments))]

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a numpy uint8 array (freed when the context closes... or never: tests only)."""
        p = C.c_void_p()
        check(lib().evqgpu_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        return arr


class Table:
    def __init__(self, ctx: Context, handle, keep=None):
        self.ctx = ctx
        self._h = handle
        self._keep = keep

    def close(self):
        if self._h:
            lib().evqgpu_table_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_rows(self) -> int:
        return lib().evqgpu_table_num_rows(self._h)

    def columns(self) -> List[dict]:
        out = []
        for i in range(lib().evqgpu_table_num_columns(self._h)):
            ci = ColumnInfo()
            check(lib().evqgpu_table_column_info(self._h, i, C.byref(ci)))
            out.append({f: (getattr(ci, f).decode() if f == "name" else getattr(ci, f)) for f, _ in ColumnInfo._fields_})
        return out

    def column(self, name: str) -> dict:
        for c in self.columns():
            if c["name"] == name:
                return c
        raise KeyError(name)

    def load(self, names: Optional[Sequence[str]] = None):
        if names is None:
            check(lib().evqgpu_table_load_columns(self._h, None, 0))
        else:
            arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
            check(lib().evqgpu_table_load_columns(self._h, arr, len(names)))
        return self

    def add_column(self, name: str, logical_type: int, encoding: int, dlevel_max: int = 0):
        rc = lib().evqgpu_table_add_column(self._h, name.encode(), logical_type, encoding, 0, dlevel_max)
        if rc < 0:
            check(-rc)

    def add_stream(self, column: str, kind: int, data: np.ndarray, bitpack_max: int = 0):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        check(lib().evqgpu_table_add_stream(self._h, column.encode(), kind, data.ctypes.data_as(C.c_void_p), data.nbytes, bitpack_max, 0))

    def set_filter(self, keep: Optional[np.ndarray]):
        """FastCSTableScan::setFilter: one bool per row (True = keep), ANDed with WHERE; None removes the filter."""
        if keep is None:
            check(lib().evqgpu_table_set_filter(self._h, None, 0, 0))
            return
        keep = np.asarray(keep, dtype=bool)
        bits = np.packbits(keep, bitorder="little")
        if bits.size == 0:
            bits = np.zeros(1, dtype=np.uint8)
        check(lib().evqgpu_table_set_filter(self._h, bits.ctypes.data_as(C.c_void_p), int(keep.size), 0))

    def read_stream(self, column: str, kind: int = P.STREAM_DATA):
        n = C.c_uint64(0)
        mx = C.c_uint32(0)
        check(lib().evqgpu_table_read_stream(self._h, column.encode(), kind, None, 0, C.byref(n), C.byref(mx)))
        out = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            check(lib().evqgpu_table_read_stream(self._h, column.encode(), kind, out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(n), C.byref(mx)))
        return out, mx.value

    def decode_column(self, name: str, row0: int = 0, nrows: Optional[int] = None) -> bytes:
        info = self.column(name)
        if nrows is None:
            nrows = self.num_rows - row0
        w = 2 if info["sql_type"] == P.BOOL else 9
        out = np.zeros(max(1, nrows * w), dtype=np.uint8)
        check(lib().evqgpu_table_decode_column(self._h, name.encode(), row0, nrows, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out[: nrows * w].tobytes()

    def decode_string_column(self, name: str, row0: int = 0, nrows: Optional[int] = None) -> bytes:
        """FastCSTableScan::fetchColumnString: the packed STRING SVector of rows [row0, row0 + nrows)."""
        if nrows is None:
            nrows = self.num_rows - row0
        need = C.c_uint64(0)
        check(lib().evqgpu_table_decode_string_column(self._h, name.encode(), row0, nrows, None, 0, C.byref(need)))
        out = np.zeros(max(1, need.value), dtype=np.uint8)
        check(lib().evqgpu_table_decode_string_column(self._h, name.encode(), row0, nrows, out.ctypes.data_as(C.c_void_p), out.nbytes,
                                                      C.byref(need)))
        return out[: need.value].tobytes()

    def get_filter(self) -> Optional[np.ndarray]:
        """The table's external row filter as one bool per row, None when it has none."""
        has = C.c_int(0)
        n = self.num_rows
        bits = np.zeros(max(1, (n + 7) // 8), dtype=np.uint8)
        check(lib().evqgpu_table_get_filter(self._h, bits.ctypes.data_as(C.c_void_p), bits.nbytes, C.byref(has)))
        if not has.value:
            return None
        return np.unpackbits(bits, bitorder="little")[:n].astype(bool)

    def write_file(self, path: str):
        check(lib().evqgpu_table_write_file(self._h, path.encode()))


class Query:
    def __init__(self, ctx: Context, plan: P.QueryPlan):
        self.ctx = ctx
        self.plan = plan
        self._pc = _PlanC(plan)
        self._h = C.c_void_p()
        check(lib().evqgpu_query_create(ctx._h, C.byref(self._pc.desc), C.byref(self._h)))
        n = lib().evqgpu_query_num_columns(self._h)
        self.types = [lib().evqgpu_query_column_type(self._h, i) for i in range(n)]

    def close(self):
        if self._h:
            lib().evqgpu_query_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _tables(self, tables):
        arr = (C.c_void_p * len(tables))(*[t._h for t in tables])
        return arr

    def execute(self, tables: Sequence[Table]):
        check(lib().evqgpu_query_execute(self._h, self._tables(tables), len(tables)))
        return self

    def merge_rows(self, rows):
        """Coordinator plans (QUERY_COORDINATOR): feed partial rows [(20-byte key, saved states)] a shard returned."""
        body = b"".join(k + d for k, d in rows)
        starts, ends, pos = [], [], 0
        for k, d in rows:
            starts.append(pos)
            pos += len(k) + len(d)
            ends.append(pos)
        n = len(rows)
        buf = C.create_string_buffer(body, max(1, len(body)))
        check(lib().evqgpu_query_merge_rows(self._h, C.cast(buf, C.c_void_p), (C.c_uint64 * max(1, n))(*starts), (C.c_uint64 * max(1, n))(*ends), n))
        return self

    def merge_finish(self):
        check(lib().evqgpu_query_merge_finish(self._h))
        return self

    def prepare(self, tables: Sequence[Table]):
        """Multi-rank jobs: COLLECTIVE - every rank calls it before enqueue() whenever its table set changes (execute() does
        it implicitly); the ranks agree on slot assignment and state layout, a local failure fails all ranks."""
        check(lib().evqgpu_query_prepare(self._h, self._tables(tables), len(tables)))
        return self

    def enqueue(self, tables: Sequence[Table]):
        check(lib().evqgpu_query_enqueue(self._h, self._tables(tables), len(tables)))
        return self

    def finish(self):
        check(lib().evqgpu_query_finish(self._h))
        return self

    def merge(self):
        check(lib().evqgpu_query_merge(self._h))
        return self

    @property
    def num_rows(self) -> int:
        n = C.c_uint64(0)
        check(lib().evqgpu_query_num_rows(self._h, C.byref(n)))
        return n.value

    def fetch_packed(self, row0: int = 0, max_rows: Optional[int] = None) -> List[bytes]:
        """Result columns in the packed SVector encoding (9 B per numeric element, 2 B per bool)."""
        if max_rows is None:
            max_rows = self.num_rows - row0
        widths = [2 if t == P.BOOL else 9 for t in self.types]
        bufs = [np.zeros(max(1, max_rows * w), dtype=np.uint8) for w in widths]
        ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data_as(C.c_void_p).value for b in bufs])
        got = C.c_uint64(0)
        check(lib().evqgpu_query_fetch(self._h, row0, max_rows, ptrs, C.byref(got)))
        return [b[: got.value * w].tobytes() for b, w in zip(bufs, widths)]

    def fetch_strings(self, column: int, row0: int = 0, max_rows: Optional[int] = None) -> bytes:
        """A string result column as packed STRING SVector elements ([u32 length][bytes][tag])."""
        if max_rows is None:
            max_rows = self.num_rows - row0
        got, need = C.c_uint64(0), C.c_uint64(0)
        check(lib().evqgpu_query_fetch_strings(self._h, column, row0, max_rows, None, 0, C.byref(got), C.byref(need)))
        buf = np.zeros(max(1, need.value), dtype=np.uint8)
        check(lib().evqgpu_query_fetch_strings(self._h, column, row0, max_rows, buf.ctypes.data_as(C.c_void_p), buf.nbytes,
                                               C.byref(got), C.byref(need)))
        return buf[: need.value].tobytes()

    def fetch_partial(self) -> List[tuple]:
        """The groups as PartialGroupByExpression rows: [(20-byte SHA-1 group key, saved states)] (plan flag QUERY_WIRE)."""
        n = self.num_rows
        keys = np.zeros(max(1, n) * 20, dtype=np.uint8)
        offs = (C.c_uint64 * (n + 1))()
        got = C.c_uint64(0)
        need = C.c_uint64(0)
        check(lib().evqgpu_query_fetch_partial(self._h, 0, n, keys.ctypes.data_as(C.c_void_p), None, 0, offs, C.byref(got), C.byref(need)))
        data = np.zeros(max(1, need.value), dtype=np.uint8)
        check(lib().evqgpu_query_fetch_partial(self._h, 0, n, keys.ctypes.data_as(C.c_void_p), data.ctypes.data_as(C.c_void_p), data.nbytes,
                                               offs, C.byref(got), C.byref(need)))
        kb, db = keys.tobytes(), data.tobytes()
        return [(kb[20 * i: 20 * i + 20], db[offs[i]: offs[i + 1]]) for i in range(got.value)]

    def store_cache(self, path: str):
        """The groups as the query cache entry the reference's partial operator stores (.qc file)."""
        check(lib().evqgpu_query_store_cache(self._h, path.encode()))

    def order_by(self, specs: Sequence[tuple]):
        """OrderByExpression over the result: [(result column, descending)], most significant first."""
        arr = (SortSpec * max(1, len(specs)))(*[SortSpec(int(c), 1 if d else 0) for c, d in specs])
        check(lib().evqgpu_query_order_by(self._h, arr, len(specs)))

    def limit(self, limit: int, offset: int = 0):
        """LimitExpression: keep result rows [offset, offset + limit)."""
        check(lib().evqgpu_query_limit(self._h, int(limit), int(offset)))

    def rows(self) -> List[tuple]:
        """Rows as python tuples (None = NULL): convenience for order-insensitive comparisons in tests."""
        cols = self.fetch_packed()
        n = len(cols[0]) // (2 if self.types[0] == P.BOOL else 9) if cols else 0
        out_cols = []
        for ci, (raw, t) in enumerate(zip(cols, self.types)):
            a = np.frombuffer(raw, dtype=np.uint8)
            if t == P.STRING:
                # fetch_packed holds the dictionary codes; the values come from evqgpu_query_fetch_strings
                buf = self.fetch_strings(ci)
                vals, pos = [], 0
                for _ in range(n):
                    ln = int.from_bytes(buf[pos:pos + 4], "little")
                    vals.append(None if buf[pos + 4 + ln] & 1 else buf[pos + 4:pos + 4 + ln])
                    pos += 5 + ln
                assert pos == len(buf)
            elif t == P.BOOL:
                a = a.reshape(n, 2)
                vals = [None if tag & 1 else bool(v) for v, tag in zip(a[:, 0].tolist(), a[:, 1].tolist())]
            else:
                a = a.reshape(n, 9)
                bits = np.ascontiguousarray(a[:, :8]).view("<u8").reshape(n)
                if t == P.FLOAT64:
                    v = bits.view(np.float64).tolist()
                elif t == P.INT64:
                    v = bits.view(np.int64).tolist()
                else:
                    v = bits.tolist()
                vals = [None if tag & 1 else x for x, tag in zip(v, a[:, 8].tolist())]
            out_cols.append(vals)
        return [tuple(c[i] for c in out_cols) for i in range(n)]

    def stats(self) -> dict:
        s = QueryStats()
        check(lib().evqgpu_query_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in QueryStats._fields_ if f != "reserved"}

    def kernel_source(self) -> str:
        return (lib().evqgpu_query_kernel_source(self._h) or b"").decode()
