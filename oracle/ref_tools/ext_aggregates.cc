/**
 * Extension aggregates min / max / mean / sum(float64) for the oracle build.
 * (TEST INFRASTRUCTURE - linked only into oracle/_ref/evqlref.)
 *
 * The reference snapshot registers only count / count_distinct / sum(int64|uint64)
 * (sql/defaults.cc:49-54); min/max/mean survive only as commented-out legacy code
 * (sql/expressions/aggregate.cc:222-441).  BASELINE.json's north_star names them,
 * so they are added here through the reference's own plugin point
 * SymbolTable::registerFunction (sql/runtime/symboltable.cc:52-62), following the
 * 0.5.0 static-typing convention of sum_uint64 (aggregate.cc:184-219):
 *
 *   min/max(T) -> T          state {T value; u64 seen}   NULL inputs are skipped
 *   mean(T)    -> float64    state {double sum; u64 n}    NULL inputs are skipped
 *   sum(float64) -> float64  state {double}              NULL adds 0.0 (like sum_uint64)
 *
 * "NULL inputs are skipped" is the legacy behaviour (aggregate.cc:246-255, :304-322);
 * a value carries STAG_NULL only when the argument is a bare column reference
 * (every pure function drops tags).
 */
#include <eventql/sql/SFunction.h>
#include <eventql/sql/svalue.h>
#include <eventql/sql/runtime/runtime.h>
#include <eventql/util/io/outputstream.h>
#include <eventql/util/io/inputstream.h>

namespace evqlref {
using namespace csql;

template <typename T> struct Pop;
template <> struct Pop<uint64_t> {
  static void pop(VMStack* s, uint64_t* v, STag* t) {
    memcpy(v, s->top, 8); memcpy(t, s->top + 8, 1); s->top += 9;
  }
  static void push(VMStack* s, uint64_t v) { pushUInt64(s, v); }
};
template <> struct Pop<int64_t> {
  static void pop(VMStack* s, int64_t* v, STag* t) {
    memcpy(v, s->top, 8); memcpy(t, s->top + 8, 1); s->top += 9;
  }
  static void push(VMStack* s, int64_t v) { pushInt64(s, v); }
};
template <> struct Pop<double> {
  static void pop(VMStack* s, double* v, STag* t) {
    memcpy(v, s->top, 8); memcpy(t, s->top + 8, 1); s->top += 9;
  }
  static void push(VMStack* s, double v) { pushFloat64(s, v); }
};

template <typename T> struct MinMaxState { T value; uint64_t seen; };

template <typename T, bool IS_MAX>
struct MinMax {
  typedef MinMaxState<T> S;
  static void acc(sql_txn*, void* self, VMStack* stack) {
    S* s = static_cast<S*>(self);
    T v; STag tag;
    Pop<T>::pop(stack, &v, &tag);
    if (tag & STAG_NULL) return;
    if (!s->seen || (IS_MAX ? v > s->value : v < s->value)) s->value = v;
    s->seen = 1;
  }
  static void get(sql_txn*, void* self, VMStack* stack) {
    Pop<T>::push(stack, static_cast<S*>(self)->value);
  }
  static void reset(sql_txn*, void* self) { memset(self, 0, sizeof(S)); }
  static void merge(sql_txn*, void* self, const void* other) {
    S* s = static_cast<S*>(self);
    const S* o = static_cast<const S*>(other);
    if (!o->seen) return;
    if (!s->seen || (IS_MAX ? o->value > s->value : o->value < s->value)) s->value = o->value;
    s->seen = 1;
  }
  static void save(sql_txn*, const void* self, OutputStream* os) { os->write((const char*) self, sizeof(S)); }
  static void load(sql_txn*, void* self, InputStream* is) { is->readNextBytes(self, sizeof(S)); }
  static SFunction fn(SType t) {
    return SFunction({ t }, t, sizeof(S), &acc, &get, &reset, &reset, nullptr, &merge, &save, &load);
  }
};

struct MeanState { double sum; uint64_t n; };

template <typename T>
struct Mean {
  static void acc(sql_txn*, void* self, VMStack* stack) {
    MeanState* s = static_cast<MeanState*>(self);
    T v; STag tag;
    Pop<T>::pop(stack, &v, &tag);
    if (tag & STAG_NULL) return;
    s->sum += (double) v;
    s->n += 1;
  }
  static void get(sql_txn*, void* self, VMStack* stack) {
    MeanState* s = static_cast<MeanState*>(self);
    pushFloat64(stack, s->sum / (double) s->n);
  }
  static void reset(sql_txn*, void* self) { memset(self, 0, sizeof(MeanState)); }
  static void merge(sql_txn*, void* self, const void* other) {
    static_cast<MeanState*>(self)->sum += static_cast<const MeanState*>(other)->sum;
    static_cast<MeanState*>(self)->n += static_cast<const MeanState*>(other)->n;
  }
  static void save(sql_txn*, const void* self, OutputStream* os) { os->write((const char*) self, sizeof(MeanState)); }
  static void load(sql_txn*, void* self, InputStream* is) { is->readNextBytes(self, sizeof(MeanState)); }
  static SFunction fn(SType t) {
    return SFunction({ t }, SType::FLOAT64, sizeof(MeanState), &acc, &get, &reset, &reset, nullptr, &merge, &save, &load);
  }
};

struct SumF64 {
  static void acc(sql_txn*, void* self, VMStack* stack) { *static_cast<double*>(self) += popFloat64(stack); }
  static void get(sql_txn*, void* self, VMStack* stack) { pushFloat64(stack, *static_cast<double*>(self)); }
  static void reset(sql_txn*, void* self) { memset(self, 0, sizeof(double)); }
  static void merge(sql_txn*, void* self, const void* other) { *static_cast<double*>(self) += *static_cast<const double*>(other); }
  static void save(sql_txn*, const void* self, OutputStream* os) { os->write((const char*) self, sizeof(double)); }
  static void load(sql_txn*, void* self, InputStream* is) { is->readNextBytes(self, sizeof(double)); }
  static SFunction fn() {
    return SFunction({ SType::FLOAT64 }, SType::FLOAT64, sizeof(double), &acc, &get, &reset, &reset, nullptr, &merge, &save, &load);
  }
};

void registerExtensionAggregates(csql::SymbolTable* sym) {
  sym->registerFunction("min", MinMax<uint64_t, false>::fn(SType::UINT64));
  sym->registerFunction("min", MinMax<int64_t, false>::fn(SType::INT64));
  sym->registerFunction("min", MinMax<double, false>::fn(SType::FLOAT64));
  sym->registerFunction("max", MinMax<uint64_t, true>::fn(SType::UINT64));
  sym->registerFunction("max", MinMax<int64_t, true>::fn(SType::INT64));
  sym->registerFunction("max", MinMax<double, true>::fn(SType::FLOAT64));
  sym->registerFunction("mean", Mean<uint64_t>::fn(SType::UINT64));
  sym->registerFunction("mean", Mean<int64_t>::fn(SType::INT64));
  sym->registerFunction("mean", Mean<double>::fn(SType::FLOAT64));
  sym->registerFunction("sum", SumF64::fn());
}

} // namespace evqlref
