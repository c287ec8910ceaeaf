"""Query plans in the wire shape of include/evqgpu.h (evqgpu_insn / evqgpu_expr / evqgpu_query_desc).

Pure Python, no native code: used by the ctypes binding (eventql_b200.capi) to hand plans to the
C ABI, and by the tests / the oracle as the common description of "the same query".

The expression model mirrors the reference's query tree after planning
(sql/qtree/{ColumnReferenceNode,LiteralExpressionNode,CallExpressionNode,IfExpressionNode}):
typing is static, implicit conversions are already explicit `to_<type>` calls, and every call
carries its resolved symbol string `name#ret/arg;arg;` (sql/runtime/symboltable.cc:33-39).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

# csql::SType (sql/svalue.h:41-49)
NIL, UINT64, INT64, FLOAT64, BOOL, STRING, TIMESTAMP64 = range(7)
TYPE_NAMES = ["nil", "uint64", "int64", "float64", "bool", "string", "timestamp64"]
TYPE_BY_NAME = {n: i for i, n in enumerate(TYPE_NAMES)}
STAG_NULL = 1   # csql::STag (sql/svalue.h:51-56)

# evqgpu_insn.op
X_CALL, X_LITERAL, X_INPUT, X_IF = 1, 3, 4, 6

QUERY_GROUPBY = 1
QUERY_PARTIAL = 2
QUERY_WIRE = 4        # groups are fetched as PartialGroupByExpression rows (Query.fetch_partial)
QUERY_COORDINATOR = 8  # the coordinator side of a cluster GROUP BY: no scan, fed with shards' partial rows (Query.merge_rows)

# cstable enums (io/cstable/cstable.h:112-130)
COL_SUBRECORD, COL_BOOLEAN, COL_UNSIGNED_INT, COL_SIGNED_INT, COL_STRING, COL_FLOAT, COL_DATETIME = range(7)
ENC_BOOLEAN_BITPACKED = 1
ENC_UINT32_BITPACKED = 10
ENC_UINT32_PLAIN = 11
ENC_UINT64_PLAIN = 12
ENC_UINT64_LEB128 = 13
ENC_FLOAT_IEEE754 = 14
ENC_STRING_PLAIN = 100
STREAM_DATA, STREAM_RLEVEL, STREAM_DLEVEL = 1, 2, 3


def symbol(name: str, ret: int, args: Sequence[int]) -> str:
    return name.lower() + "#" + TYPE_NAMES[ret] + "/" + "".join(TYPE_NAMES[a] + ";" for a in args)


class Expr:
    type: int

    # convenience operators for hand-built plans in tests
    def _bin(self, name, other):
        return call(name, self, lit(other) if not isinstance(other, Expr) else other)

    def __add__(self, o): return self._bin("add", o)
    def __sub__(self, o): return self._bin("sub", o)
    def __mul__(self, o): return self._bin("mul", o)
    def __truediv__(self, o): return self._bin("div", o)
    def __mod__(self, o): return self._bin("mod", o)
    def __lt__(self, o): return self._bin("lt", o)
    def __le__(self, o): return self._bin("lte", o)
    def __gt__(self, o): return self._bin("gt", o)
    def __ge__(self, o): return self._bin("gte", o)
    def eq(self, o): return self._bin("eq", o)
    def neq(self, o): return self._bin("neq", o)
    def __and__(self, o): return self._bin("logical_and", o)
    def __or__(self, o): return self._bin("logical_or", o)
    def __invert__(self): return call("neg", self)


@dataclass(eq=False)
class Col(Expr):
    index: int
    type: int


@dataclass(eq=False)
class Lit(Expr):
    value: Union[int, float, bool, str]
    type: int


@dataclass(eq=False)
class Call(Expr):
    symbol: str
    args: List[Expr]
    type: int

    @property
    def name(self) -> str:
        return self.symbol.split("#", 1)[0]


@dataclass(eq=False)
class If(Expr):
    cond: Expr
    then: Expr
    otherwise: Expr

    @property
    def type(self) -> int:  # type: ignore[override]
        return self.then.type


def lit(v, type: Optional[int] = None) -> Lit:
    """Literal typing follows runtime/queryplanbuilder.cc:1520-1530: no '-' -> UINT64, '-' -> INT64, '.' -> FLOAT64."""
    if type is None:
        if isinstance(v, bool):
            type = BOOL
        elif isinstance(v, int):
            type = UINT64 if v >= 0 else INT64
        elif isinstance(v, float):
            type = FLOAT64
        elif isinstance(v, str):
            type = STRING
        else:
            raise TypeError(v)
    return Lit(v, type)


# name -> [(arg types, return type, allow_arg_conversion, is_aggregate)] in the registration order of
# sql/defaults.cc:38-171 (+ the typed extension aggregates of oracle/ref_tools/ext_aggregates.cc)
def _registry():
    R = {}

    def reg(name, args, ret, conv=True, agg=False):
        R.setdefault(name, []).append((tuple(args), ret, conv, agg))

    reg("count", [NIL], UINT64, agg=True)
    reg("sum", [INT64], INT64, agg=True)
    reg("sum", [UINT64], UINT64, agg=True)
    reg("count_distinct", [UINT64], UINT64, agg=True)      # sql/defaults.cc:50, aggregate.cc:80-137
    reg("logical_and", [BOOL, BOOL], BOOL)
    reg("logical_or", [BOOL, BOOL], BOOL)
    reg("neg", [BOOL], BOOL)
    for t in (UINT64, INT64, FLOAT64, TIMESTAMP64):
        reg("cmp", [t, t], INT64)
    for name in ("eq", "neq"):
        for t in (UINT64, INT64, FLOAT64, BOOL, STRING, TIMESTAMP64):   # sql/defaults.cc:65-76
            reg(name, [t, t], BOOL, conv=False)
    for name in ("lt", "lte", "gt", "gte"):
        for t in (UINT64, INT64, FLOAT64, STRING, TIMESTAMP64):         # sql/defaults.cc:77-96

            # boolean.cc:415-430: lt_int64 is the one comparison that allows argument conversion
            reg(name, [t, t], BOOL, conv=(name == "lt" and t == INT64))
    for t in (UINT64, INT64, FLOAT64, BOOL, TIMESTAMP64):
        reg("to_nil", [t], NIL)
    for t in (UINT64, FLOAT64, BOOL, TIMESTAMP64):
        reg("to_int64", [t], INT64)
    reg("to_timestamp64", [INT64], TIMESTAMP64)
    reg("to_timestamp64", [FLOAT64], TIMESTAMP64)
    reg("from_timestamp", [INT64], TIMESTAMP64)
    reg("from_timestamp", [FLOAT64], TIMESTAMP64)
    reg("date_trunc", [STRING, TIMESTAMP64], TIMESTAMP64)
    reg("startswith", [STRING, STRING], BOOL)                           # sql/defaults.cc:149-150, expressions/string.cc:52-74
    reg("endswith", [STRING, STRING], BOOL)
    for name in ("add", "sub", "mul", "div", "mod", "pow"):
        for t in (UINT64, INT64, FLOAT64):
            reg(name, [t, t], t)
    # extension aggregates
    for name in ("min", "max"):
        for t in (UINT64, INT64, FLOAT64):
            reg(name, [t], t, agg=True)
    for t in (UINT64, INT64, FLOAT64):
        reg("mean", [t], FLOAT64, agg=True)
    reg("sum", [FLOAT64], FLOAT64, agg=True)
    return R


REGISTRY = _registry()
# implicit conversions (sql/defaults.cc:40-46)
IMPLICIT = {(UINT64, INT64)} | {(t, NIL) for t in (UINT64, INT64, FLOAT64, BOOL, STRING, TIMESTAMP64)}

AGGREGATE_NAMES = {n for n, sigs in REGISTRY.items() if any(s[3] for s in sigs)}


def is_aggregate_symbol(sym: str) -> bool:
    return sym.split("#", 1)[0] in AGGREGATE_NAMES


def call(name: str, *args: Expr) -> Call:
    """SymbolTable::resolve (runtime/symboltable.cc:71-160) + CallExpressionNode::newNode
    (qtree/CallExpressionNode.cc:32-101): exact match first, then the first candidate reachable by
    implicit conversions, which are materialised as to_<type> calls."""
    name = name.lower()
    args = [a if isinstance(a, Expr) else lit(a) for a in args]
    cands = REGISTRY.get(name)
    if not cands:
        raise KeyError("method not found: %s" % name)
    at = tuple(a.type for a in args)
    match = None
    for sig in cands:
        if sig[0] == at:
            match = sig
            break
    if match is None:
        for sig in cands:
            if len(sig[0]) != len(at) or not sig[2]:
                continue
            if all(a == b or (a, b) in IMPLICIT for a, b in zip(at, sig[0])):
                match = sig
                break
    if match is None:
        raise TypeError("type error for %s<%s>" % (name, ", ".join(TYPE_NAMES[t] for t in at)))
    conv = []
    for a, want in zip(args, match[0]):
        conv.append(a if a.type == want else call("to_" + TYPE_NAMES[want], a))
    return Call(symbol(name, match[1], match[0]), conv, match[1])


def find_aggregate(e: Expr) -> Optional[Call]:
    """QueryTreeUtil::findAggregateExpression (qtree/QueryTreeUtil.cc:209-224): first aggregate call, depth first."""
    if isinstance(e, Call):
        if is_aggregate_symbol(e.symbol):
            return e
        for a in e.args:
            r = find_aggregate(a)
            if r is not None:
                return r
    elif isinstance(e, If):
        for a in (e.cond, e.then, e.otherwise):
            r = find_aggregate(a)
            if r is not None:
                return r
    return None


def _imm(value, type_: int) -> int:
    if type_ == FLOAT64:
        return struct.unpack("<Q", struct.pack("<d", float(value)))[0]
    if type_ == BOOL:
        return 1 if value else 0
    if type_ == INT64:
        return int(value) & 0xFFFFFFFFFFFFFFFF
    if type_ == NIL:
        return 0
    return int(value) & 0xFFFFFFFFFFFFFFFF


@dataclass
class Program:
    """Flattened postfix form: list of (op, type, nargs, arg, imm) + string pool."""
    insns: List[Tuple[int, int, int, int, int]] = field(default_factory=list)
    strings: bytes = b""


def flatten(e: Optional[Expr], fn_id=None) -> Program:
    """Postfix serialisation. `fn_id(symbol) -> int` supplies evqgpu function ids (capi passes
    evqgpu_function_lookup); when None the symbol index into prog.symbols is stored instead."""
    p = Program()
    if e is None:
        return p
    symbols: List[str] = []

    def emit(n: Expr):
        if isinstance(n, Col):
            p.insns.append((X_INPUT, n.type, 0, n.index, 0))
        elif isinstance(n, Lit):
            if n.type == STRING:
                raw = n.value.encode() if isinstance(n.value, str) else bytes(n.value)
                off = len(p.strings)
                p.strings += raw
                p.insns.append((X_LITERAL, STRING, 0, 0, (off << 32) | len(raw)))
            else:
                p.insns.append((X_LITERAL, n.type, 0, 0, _imm(n.value, n.type)))
        elif isinstance(n, Call):
            for a in n.args:
                emit(a)
            if fn_id is not None:
                fid = fn_id(n.symbol)
                if fid < 0:
                    raise NotImplementedError("function not available on the device path: %s" % n.symbol)
            else:
                if n.symbol not in symbols:
                    symbols.append(n.symbol)
                fid = symbols.index(n.symbol)
            p.insns.append((X_CALL, n.type, len(n.args), fid, 0))
        elif isinstance(n, If):
            emit(n.cond)
            emit(n.then)
            emit(n.otherwise)
            p.insns.append((X_IF, n.type, 3, 0, 0))
        else:
            raise TypeError(n)

    emit(e)
    p.symbols = symbols  # type: ignore[attr-defined]
    return p


@dataclass
class QueryPlan:
    """evqgpu_query_desc: a fused FastCSTableScan (+ GroupByExpression)."""
    input_columns: List[str]
    select: List[Expr]
    where: Optional[Expr] = None
    group: List[Expr] = field(default_factory=list)
    flags: int = QUERY_GROUPBY
    expected_groups: int = 0

    @property
    def is_groupby(self) -> bool:
        return bool(self.flags & QUERY_GROUPBY)
