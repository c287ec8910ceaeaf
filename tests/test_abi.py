"""The C-ABI library loads and exports every symbol include/evqgpu.h declares (no GPU needed).

Also the device-free half of the product: plan intake, kernel text generation and NVRTC compilation
for sm_100a of every kernel variant (evqgpu_debug_generate), and the "fails loudly" contract.
"""
import ctypes
import os
import re
import subprocess

import pytest

from eventql_b200 import capi, plan as P
from tests import common as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "evqgpu.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"EVQGPU_API\s+[^;(]*?\b(evqgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_binding_lists():
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(native_lib):
    for sym in header_symbols():
        assert hasattr(native_lib, sym), sym
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], stdout=subprocess.PIPE, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (evqgpu_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    # nothing else leaks out of the library: the ABI is exactly the header
    assert exported == set(header_symbols())


def test_abi_version(native_lib):
    m = re.search(r"#define EVQGPU_ABI_VERSION (\d+)", open(HEADER).read())
    assert native_lib.evqgpu_abi_version() == int(m.group(1))


def test_no_torch_types_in_header():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)   # declarations only
    assert "torch" not in text.lower() and "at::" not in text and "std::" not in text and "Tensor" not in text


def test_function_registry_matches_plan_module(native_lib):
    """Every symbol the plan module can resolve is known to the device path, with the same aggregate flag."""
    missing = []
    for name, sigs in P.REGISTRY.items():
        for args, ret, _conv, agg in sigs:
            sym = P.symbol(name, ret, args)
            fid = native_lib.evqgpu_function_lookup(sym.encode())
            if fid < 0:
                missing.append(sym)
                continue
            assert native_lib.evqgpu_function_symbol(fid).decode() == sym
            assert bool(native_lib.evqgpu_function_is_aggregate(fid)) == agg, sym
    # pow / cmp helpers and the ordering comparisons of strings (dictionary codes carry no order) are not part of the device
    # path; everything the named configs need is
    allowed_missing = {s for s in missing if s.split("#")[0] in ("cmp", "pow", "from_timestamp") or
                       (s.split("#")[0] in ("lt", "lte", "gt", "gte") and "string" in s)}
    assert set(missing) == allowed_missing, sorted(set(missing) - allowed_missing)
    assert native_lib.evqgpu_function_lookup(b"no_such_fn#uint64/uint64;") == -1
    assert native_lib.evqgpu_function_symbol(100000) is None


# column statistics (value bits, longest LEB128 value) as the loader computes them for the synthetic tables
STATS = {"shipdate": (14, 2), "discount": (7, 1), "quantity": (7, 1), "price": (28, 4), "tax": (7, 1), "flag": (7, 1),
         "status": (7, 1), "ekey": (64, 0), "v": (21, 3), "time": (56, 8), "sensor_id": (14, 2), "value": (21, 3)}


def _cols(spec, stats=True):
    return [(T.sql_type_of(s), s["encoding"], 1 if s.get("null_every") else 0) + (STATS.get(s["name"], (0, 0)) if stats else (0, 0))
            for s in spec]


@pytest.mark.parametrize("case", ["q6_leb", "q6_plain", "q1_dense", "q1_nostats", "q1_null", "hc_hash", "ts_hash", "scan_only_mixed",
                                  "scan_only_required"])
def test_kernel_text_compiles_for_sm100a(native_lib, case):
    """Codegen + NVRTC (--gpu-architecture=sm_100a) of every kernel tier, without a device."""
    if case == "q6_leb":
        spec = T.lineitem_spec()
        _, plan = T.q6(spec)
        tier, slots = 1, 1
    elif case == "q6_plain":
        spec = T.lineitem_spec(P.ENC_UINT64_PLAIN)
        _, plan = T.q6(spec)
        tier, slots = 1, 1
    elif case == "q1_dense":
        spec = T.lineitem_spec()
        _, plan = T.q1(spec)
        tier, slots = 1, 4
    elif case == "q1_nostats":
        spec = T.lineitem_spec()
        _, plan = T.q1(spec)
        tier, slots = 1, 4
    elif case == "q1_null":
        spec = T.lineitem_spec(null_every=7)
        _, plan = T.q1(spec)
        tier, slots = 1, 9
    elif case == "hc_hash":
        spec = T.events_spec()
        _, plan = T.q_highcard(spec)
        tier, slots = 2, 0
    elif case == "ts_hash":
        spec = T.readings_spec(0)
        _, plan = T.q_timeseries(spec)
        tier, slots = 2, 0
    else:
        spec = T.mixed_spec(0 if case == "scan_only_required" else 7)
        c, names = T.cols_of(spec)
        plan = P.QueryPlan(names, [c["a"], c["f"], c["bo"], c["c"] + c["d"], c["big"]], where=c["b"] < 50, flags=0)
        tier, slots = 0, 0
    src, cubin_bytes = capi.debug_generate(plan, _cols(spec, case != "q1_nostats"), tier=tier, dense_slots=slots, compile=True)
    assert "evq_scan" in src and cubin_bytes > 1000
    # required and flat optional columns take the fast kernel (consecutive rows per thread); optional ones read their
    # presence bits from the level stream
    nullable = any(s.get("null_every") for s in spec)
    assert "evq_fast_decode" in src
    assert ("cols.n" in src and "evq_fast_presence<" in src.split("void evq_fast_nulls(", 1)[1].split("evq_fast_decode(", 1)[0]) == nullable
    # the TMA bulk copy + mbarrier pipeline is part of every variant
    assert "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes" in src


def test_unsupported_plans_fail_loudly(native_lib):
    spec = T.lineitem_spec()
    c, names = T.cols_of(spec)
    # string group key: outside the numeric device path -> EVQGPU_ERR_UNSUPPORTED, never a CPU fallback
    plan = P.QueryPlan(names, [P.call("count", P.lit(1))], group=[P.lit("x")])
    with pytest.raises(capi.EvqError) as ei:
        capi.debug_generate(plan, _cols(spec), compile=False)
    assert ei.value.status == 2
    # aggregate in WHERE -> malformed plan
    plan = P.QueryPlan(names, [P.call("count", P.lit(1))], where=P.call("sum", c["price"]) > 0)
    with pytest.raises(capi.EvqError) as ei:
        capi.debug_generate(plan, _cols(spec), compile=False)
    assert ei.value.status == 1


def test_first_row_select_item_compiles_with_one_128_bit_cas(native_lib):
    """A scalar select item that is not a function of the group key takes the value of the group's first row
    (groupby.cc:161-172): (row ordinal | tag, value) pairs updated with atom.cas.b128, in the dense and the hash tier."""
    spec = T.lineitem_spec()
    c, names = T.cols_of(spec)
    plan = P.QueryPlan(names, [c["price"], P.call("count", P.lit(1))], group=[c["flag"]])
    for tier, slots in ((1, 2), (2, 0)):
        src, cubin_bytes = capi.debug_generate(plan, _cols(spec), tier=tier, dense_slots=slots, compile=True)
        assert "evq_first_update(EVQ_GPTR(" in src and "atom.global.relaxed.gpu.cas.b128" in src and cubin_bytes > 1000


def test_partitioned_hash_tier_kernels_compile(native_lib):
    """The hash tier beyond L2 (strategy 4): pass 1 with per-partition shared-memory bins inside evq_scan, evq_repart, and
    evq_agg_smem with TMA bulk copies of the table slices - NVRTC for sm_100a.  Narrow columns pack into one record word."""
    spec = T.events_spec()
    _, plan = T.q_highcard(spec)
    for cols, nrec in ((_cols(spec), 2),                                                    # full-range 64-bit keys: two words
                       ([(P.UINT64, P.ENC_UINT64_PLAIN, 0, 24, 0, 0, 0, 10_000_000),
                         (P.UINT64, P.ENC_UINT64_LEB128, 0, 20, 3, 0, 0, 999_999)], 1)):      # 24-bit keys + 20-bit values: one
        src, cubin_bytes = capi.debug_generate(plan, cols, tier=4, compile=True)
        assert "#define EVQ_NREC %d\n" % nrec in src, [l for l in src.splitlines() if "EVQ_NREC" in l][:2]
        for name in ("evq_repart", "evq_agg_smem", "evq_agg_part"):
            assert "void __launch_bounds__" in src and name + "(" in src
        assert "cp.async.bulk.global.shared::cta.bulk_group" in src and "atom.shared.add.u32" in src and cubin_bytes > 1000


def test_context_without_device_raises(native_lib):
    """The product path has no host execution mode: no CUDA device -> hard error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(capi.EvqError) as ei:
        capi.Context(0)
    assert ei.value.status == 3
    assert "no CUDA device" in ei.value.message


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "eventql_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")) and "_build" not in dirpath:
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "oracle/ref_tools/ext_aggregates.cc", ""), os.path.join(dirpath, f)
    assert "oracle" not in open(HEADER).read().replace("oracle/ref_tools/ext_aggregates.cc", "")


def test_partial_cache_entry_equals_the_reference_file(native_lib):
    """The query cache entry of a partial aggregation (sql/statements/select/groupby.cc:411-432, query_cache.cc:58-75):
    tests/golden/ref_partial.json holds, per case, the .qc file the reference's own PartialGroupByExpression::execute stored
    (make_golden_partial.py, evqlref sql -P -C) next to its rows.  Encoding the reference's rows, in the file's group
    order, reproduces the file byte for byte; the file name is the reference's cache key."""
    import hashlib
    import json
    import os
    with open(os.path.join(T.ROOT, "tests", "golden", "ref_partial.json")) as fh:
        cases = json.load(fh)["cases"]
    for name, g in cases.items():
        qc = bytes.fromhex(g["qc"])
        rows = {bytes.fromhex(k): bytes.fromhex(d) for k, d in g["rows"]}
        assert qc[0] == 1 and int.from_bytes(qc[1:9], "little") == len(rows), name
        ordered, pos = [], 9
        while pos < len(qc):                      # the groups in the file's (hash map) order
            key = qc[pos:pos + 20]
            data = rows[key]
            assert qc[pos + 20:pos + 20 + len(data)] == data, name
            ordered.append((key, data))
            pos += 20 + len(data)
        assert len(ordered) == len(rows), name
        assert capi.partial_cache_encode(ordered) == qc, name
        # evqlref keys its scan with SHA1("evqlref-input") and the operator with SHA1("evqlref") (oracle/ref_tools/evqlref.cc)
        fn = capi.partial_cache_filename(hashlib.sha1(b"evqlref-input").digest(), hashlib.sha1(b"evqlref").digest())
        assert fn == g["qc_file"], name
    assert capi.partial_cache_encode([]) == bytes([1]) + bytes(8)


def test_partial_result_frames_equal_the_reference(native_lib):
    """QUERY_PARTIALAGGR_RESULT frames (transport/native/ops/query_partialaggr.cc:83-124): ref_partial.json holds the frames
    the server op's loop produces over the reference's QueryPartialAggrResultFrame for every case, unsplit and split by a
    4 KiB soft maximum (evqlref sql -P -F).  Encoding the reference's rows in frame order reproduces them byte for byte."""
    import json
    import os
    import struct
    with open(os.path.join(T.ROOT, "tests", "golden", "ref_partial.json")) as fh:
        cases = json.load(fh)["cases"]
    split_seen = False
    for name, g in cases.items():
        rows = {bytes.fromhex(k): bytes.fromhex(d) for k, d in g["rows"]}
        for soft_max, hexed in g["frames"].items():
            ref = bytes.fromhex(hexed)
            ordered, pos, nframes = [], 0, 0
            while pos < len(ref):
                op, fl, ln = struct.unpack(">HHI", ref[pos:pos + 8])
                assert op == 0x0102 and fl == (1 if pos + 8 + ln == len(ref) else 0), name
                p = pos + 8
                assert ref[p] == 0
                p += 1
                nrows, shift = 0, 0
                while True:
                    b = ref[p]
                    p += 1
                    nrows |= (b & 0x7f) << shift
                    shift += 7
                    if not b & 0x80:
                        break
                for _ in range(nrows):
                    key = ref[p:p + 20]
                    ordered.append((key, rows[key]))
                    p += 20 + len(rows[key])
                assert p == pos + 8 + ln, name
                pos = p
                nframes += 1
            split_seen = split_seen or nframes > 2
            assert len(ordered) == len(rows), name
            assert capi.partial_frames_encode(ordered, int(soft_max)) == ref, (name, soft_max)
    assert split_seen
    # no groups: one empty frame that ends the request
    assert capi.partial_frames_encode([]) == bytes([1, 2, 0, 1, 0, 0, 0, 2, 0, 0])


def test_partial_rows_are_read_back_with_the_plan(native_lib):
    """The coordinator's side of the partial-aggregation formats: the .qc entries and result frames the REFERENCE produced
    (ref_partial.json) are split into (key, saved states) rows by walking them with the plan, like GroupByMergeExpression /
    the cache load do (groupby.cc:553-615, :262-292); the rows are the reference's rows.  Truncated input is refused."""
    import json
    import os
    with open(os.path.join(T.ROOT, "tests", "golden", "ref_partial.json")) as fh:
        cases = json.load(fh)["cases"]
    plans = {name: plan for name, _sql, plan in T.partial_cases()}
    for name, g in cases.items():
        want = sorted((bytes.fromhex(k), bytes.fromhex(d)) for k, d in g["rows"])
        qc = bytes.fromhex(g["qc"])
        assert sorted(capi.partial_cache_decode(plans[name], qc)) == want, name
        assert sorted(capi.partial_rows_split(plans[name], qc[9:])) == want, name
        for soft_max, hexed in g["frames"].items():
            rows, nframes, eor = capi.partial_frames_decode(plans[name], bytes.fromhex(hexed))
            assert sorted(rows) == want and eor and nframes >= 1, (name, soft_max)
        # round trip through our own encoders
        assert capi.partial_cache_decode(plans[name], capi.partial_cache_encode(want)) == want
        rows, nframes, eor = capi.partial_frames_decode(plans[name], capi.partial_frames_encode(want, 512))
        assert rows == want and eor
        if len(qc) > 40:
            with pytest.raises(capi.EvqError) as ei:
                capi.partial_cache_decode(plans[name], qc[:-3])
            assert ei.value.status == 5
            with pytest.raises(capi.EvqError):
                capi.partial_cache_decode(plans[name], bytes([2]) + qc[1:])
