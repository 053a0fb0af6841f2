// pb_host.cpp — the host planning layer above the C ABI, class for class after the reference
// (in a Pinot server these steps run in Java; see include/pinot_b200_host.h and INTEGRATION.md).
//
//   Dictionary                     SEGL/segment/index/readers/BaseImmutableDictionary.java:45-260
//   PredicateEvaluatorProvider     CTR/operator/filter/predicate/PredicateEvaluatorProvider.java:45-95
//     RANGE  (sorted dictionary)   …/RangePredicateEvaluatorFactory.java:119-169   raw: :326-400
//     EQ / NOT_EQ                  …/EqualsPredicateEvaluatorFactory.java:83-110, NotEqualsPredicateEvaluatorFactory.java:83-110
//     IN / NOT_IN                  …/InPredicateEvaluatorFactory.java:158-188, NotInPredicateEvaluatorFactory.java:155-172
//   FilterOperatorUtils            CTR/operator/filter/FilterOperatorUtils.java:74-252
//   FilterPlanNode                 CTR/plan/FilterPlanNode.java:195-320
//   SortedIndexBasedFilterOperator CTR/operator/filter/SortedIndexBasedFilterOperator.java:53-131
//   B200PlanMaker eligibility      SURVEY.md §8b (override of InstancePlanMakerImplV2.makeSegmentPlanNode :275-294)
#include "../../../include/pinot_b200_host.h"
#include "../pb_internal.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <limits>
#include <string>
#include <vector>

namespace pinot_b200 {

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }

struct BadQuery { std::string msg; };   // BadQueryRequestException

// ---------------------------------------------------------------------------------------------- Dictionary
class Dictionary {
 public:
  explicit Dictionary(const PbColumnView& c) : _c(c) {}
  int length() const { return _c.card; }
  long long getLong(int id) const { return _c.type == PB_INT ? (long long)(int32_t)be32(_c.dict + 4ull * id) : (long long)be64(_c.dict + 8ull * id); }
  double getDouble(int id) const {
    if (_c.type == PB_FLOAT) { uint32_t u = be32(_c.dict + 4ull * id); float f; memcpy(&f, &u, 4); return f; }
    uint64_t u = be64(_c.dict + 8ull * id); double d; memcpy(&d, &u, 8); return d;
  }
  // Arrays.binarySearch convention: index if found, else -(insertionPoint + 1)
  int insertionIndexOf(const std::string& v) const {
    switch (_c.type) {
      case PB_INT: case PB_LONG: {
        long long iv; double dv;
        if (parseIntegral(v, &iv)) return search([&](int m) { long long x = getLong(m); return (x > iv) - (x < iv); });
        if (v.find_first_not_of("+-0123456789") == std::string::npos)      // an integer that does not fit a long: Long.parseLong throws
          throw BadQuery{"integral literal out of range: '" + v + "'"};
        dv = parseDouble(v);   // fractional literal against an integral dictionary: never equal
        return search([&](int m) { double x = (double)getLong(m); return (x > dv) - (x < dv); });
      }
      case PB_FLOAT: { float fv = (float)parseDouble(v); return search([&](int m) { float x = (float)getDouble(m); return (x > fv) - (x < fv); }); }
      case PB_DOUBLE: { double dv = parseDouble(v); return search([&](int m) { double x = getDouble(m); return (x > dv) - (x < dv); }); }
      default:
        return search([&](int m) {
          const uint8_t* e = _c.dict + (size_t)m * _c.entry_bytes;
          size_t el = 0; while (el < (size_t)_c.entry_bytes && e[el]) el++;
          return compareUtf8(e, el, (const uint8_t*)v.data(), v.size());
        });
    }
  }
  int indexOf(const std::string& v) const { int i = insertionIndexOf(v); return i >= 0 ? i : -1; }

  // ValueReaderComparisons.compareUtf8Bytes (SEGL/io/util/ValueReaderComparisons.java:68-139): the dictionary is sorted by
  // String.compareTo, i.e. by UTF-16 code units -- find the first byte that differs, step back to the start of its UTF-8
  // sequence, decode both sides, compare the units.  Differs from byte order only between a supplementary character (a
  // surrogate pair, 0xD800-0xDFFF) and a BMP character at or above U+E000; a prefix sorts first.
  static void utf16UnitsAt(const uint8_t* p, size_t avail, uint32_t* u1, uint32_t* u2) {
    *u1 = 0xfffd; *u2 = 0xfffd;
    if (avail == 0) { *u1 = 0; return; }
    auto cont = [&](size_t k) -> uint32_t { return k < avail ? (p[k] & 0x3Fu) : 0u; };
    const uint8_t b = p[0];
    if (b < 0x80) *u1 = b;
    else if ((b & 0xF0) < 0xE0) *u1 = ((uint32_t)(b & 0x1F) << 6) | cont(1);
    else if ((b & 0xF0) == 0xE0) *u1 = ((uint32_t)(b & 0x0F) << 12) | (cont(1) << 6) | cont(2);
    else {
      const uint32_t cp = ((uint32_t)(b & 0x07) << 18) | (cont(1) << 12) | (cont(2) << 6) | cont(3);
      if (cp >= 0x10000 && cp <= 0x10FFFF) { *u1 = 0xD800 + ((cp - 0x10000) >> 10); *u2 = 0xDC00 + ((cp - 0x10000) & 0x3FF); }
    }
  }
  static int compareUtf8(const uint8_t* a, size_t alen, const uint8_t* b, size_t blen) {
    const size_t m = std::min(alen, blen);
    size_t i = 0;
    while (i < m && a[i] == b[i]) i++;
    if (i == m) return (alen > blen) - (alen < blen);
    while (i > 0 && (b[i] & 0xC0) == 0x80) i--;
    uint32_t a1, a2, b1, b2;
    utf16UnitsAt(a + i, alen - i, &a1, &a2);
    utf16UnitsAt(b + i, blen - i, &b1, &b2);
    if (a1 != b1) return a1 < b1 ? -1 : 1;
    return (a2 > b2) - (a2 < b2);
  }

  static bool parseIntegral(const std::string& s, long long* out) {
    errno = 0; char* end = nullptr;
    long long v = strtoll(s.c_str(), &end, 10);
    if (errno || end == s.c_str() || *end) return false;
    *out = v; return true;
  }
  static double parseDouble(const std::string& s) {
    errno = 0; char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) throw BadQuery{"cannot parse numeric literal '" + s + "'"};
    return v;
  }

 private:
  template <class Cmp> int search(Cmp cmp) const {
    int lo = 0, hi = _c.card - 1;
    while (lo <= hi) {
      int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
      int r = cmp(mid);
      if (r < 0) lo = mid + 1; else if (r > 0) hi = mid - 1; else return mid;
    }
    return -(lo + 1);
  }
  const PbColumnView& _c;
};

// ---------------------------------------------------------------------------------------------- PredicateEvaluator
struct PredicateEvaluator {
  int type = 0;               // PBH_*
  int column = -1;
  bool dictionaryBased = false;
  bool alwaysTrue = false, alwaysFalse = false;
  bool exclusive = false;     // NOT_EQ / NOT_IN
  // dictionary based
  int startDictId = 0, endDictId = 0;          // RANGE: [start, end)
  std::vector<int32_t> dictIds;                // EQ/IN: matching; NOT_EQ/NOT_IN: non-matching; sorted, unique
  // raw value based
  long long ilo = 0, ihi = 0; double dlo = 0, dhi = 0; bool dloIncl = false, dhiIncl = false;
  std::vector<int64_t> rawValues;
  bool isRange() const { return type == PBH_RANGE; }
};

class PredicateEvaluatorProvider {
 public:
  static PredicateEvaluator getPredicateEvaluator(const pbh_predicate& p, const PbColumnView& c, int columnIndex) {
    PredicateEvaluator e;
    e.type = p.type; e.column = columnIndex; e.dictionaryBased = c.has_dict; e.exclusive = p.type == PBH_NOT_EQ || p.type == PBH_NOT_IN;
    if (c.has_dict) {
      Dictionary dict(c);
      const int card = dict.length();
      if (p.type == PBH_RANGE) {
        if (!p.lower) e.startDictId = 0;
        else { int ii = dict.insertionIndexOf(p.lower); e.startDictId = ii < 0 ? -(ii + 1) : (p.lower_inclusive ? ii : ii + 1); }
        if (!p.upper) e.endDictId = card;
        else { int ii = dict.insertionIndexOf(p.upper); e.endDictId = ii < 0 ? -(ii + 1) : (p.upper_inclusive ? ii + 1 : ii); }
        int n = std::max(e.endDictId - e.startDictId, 0);
        if (n == 0) e.alwaysFalse = true; else if (n == card) e.alwaysTrue = true;
      } else {
        for (int i = 0; i < p.num_values; i++) { int id = dict.indexOf(p.values[i]); if (id >= 0) e.dictIds.push_back(id); }
        std::sort(e.dictIds.begin(), e.dictIds.end());
        e.dictIds.erase(std::unique(e.dictIds.begin(), e.dictIds.end()), e.dictIds.end());
        const int n = (int)e.dictIds.size();
        if (!e.exclusive) { if (n == 0) e.alwaysFalse = true; else if (n == card) e.alwaysTrue = true; }
        else { if (n == 0) e.alwaysTrue = true; else if (n == card) e.alwaysFalse = true; }
      }
      return e;
    }
    // raw value based
    if (c.type == PB_STRING) throw BadQuery{"raw STRING predicate on column " + c.name};
    const bool integral = c.type == PB_INT || c.type == PB_LONG;
    if (p.type == PBH_RANGE) {
      if (integral) {
        const long long tmin = c.type == PB_INT ? INT32_MIN : INT64_MIN, tmax = c.type == PB_INT ? INT32_MAX : INT64_MAX;
        auto parse = [&](const char* s) { long long v; if (!Dictionary::parseIntegral(s, &v)) throw BadQuery{std::string("cannot parse integral literal '") + s + "'"}; return v; };
        e.ilo = p.lower ? parse(p.lower) : tmin; e.ihi = p.upper ? parse(p.upper) : tmax;
        if (p.lower && !p.lower_inclusive) { if (e.ilo == tmax) e.alwaysFalse = true; else e.ilo++; }   // RangePredicateEvaluatorFactory.java:334-345
        if (p.upper && !p.upper_inclusive) { if (e.ihi == tmin) e.alwaysFalse = true; else e.ihi--; }
        if (e.ilo > e.ihi) e.alwaysFalse = true;
      } else {
        e.dlo = p.lower ? Dictionary::parseDouble(p.lower) : -INFINITY; e.dhi = p.upper ? Dictionary::parseDouble(p.upper) : INFINITY;
        if (c.type == PB_FLOAT) { e.dlo = (double)(float)e.dlo; e.dhi = (double)(float)e.dhi; }
        e.dloIncl = !p.lower || p.lower_inclusive; e.dhiIncl = !p.upper || p.upper_inclusive;
      }
    } else {
      for (int i = 0; i < p.num_values; i++) {
        if (integral) { long long v; if (!Dictionary::parseIntegral(p.values[i], &v)) throw BadQuery{std::string("cannot parse integral literal '") + p.values[i] + "'"}; e.rawValues.push_back(v); }
        else {
          // The device compares IEEE-754 bit patterns.  That IS the reference for IN / NOT_IN (fastutil DoubleSet / FloatSet
          // compare Double.doubleToLongBits: -0.0 is not in {0.0}; InPredicateEvaluatorFactory.java:341-362), but EQ / NOT_EQ
          // compare with == / != (EqualsPredicateEvaluatorFactory.java:336-337): 0.0 and -0.0 are equal, NaN equals nothing.
          double d = Dictionary::parseDouble(p.values[i]);
          if (c.type == PB_FLOAT) d = (double)(float)d;
          const bool eq = p.type == PBH_EQ || p.type == PBH_NOT_EQ;
          if (eq && d != d) continue;                               // x = NaN matches nothing, x <> NaN everything
          if (d != d) d = std::numeric_limits<double>::quiet_NaN(); // doubleToLongBits collapses every NaN to the canonical one
          int64_t b; memcpy(&b, &d, 8); e.rawValues.push_back(b);
          if (eq && d == 0.0) { const double z = -d; memcpy(&b, &z, 8); e.rawValues.push_back(b); }   // the other zero
        }
      }
    }
    return e;
  }
};

// ---------------------------------------------------------------------------------------------- filter operators
enum OpKind { OP_EMPTY, OP_MATCH_ALL, OP_SORTED, OP_INVERTED, OP_SCAN, OP_AND, OP_OR, OP_NOT, OP_BITMAP };
struct FilterOperator {
  OpKind kind = OP_EMPTY;
  PredicateEvaluator ev;
  std::vector<int32_t> docIdRanges;   // OP_SORTED: inclusive (start,end) pairs
  const uint8_t* bitmap = nullptr;    // OP_BITMAP (BitmapBasedFilterOperator): serialized RoaringBitmap, flipped over [0, numDocs) when exclusive
  uint64_t bitmapLen = 0;
  bool bitmapExclusive = false;
  int bitmapColumn = -1;
  std::vector<std::unique_ptr<FilterOperator>> children;
};
using OpPtr = std::unique_ptr<FilterOperator>;
static OpPtr mk(OpKind k) { OpPtr p(new FilterOperator()); p->kind = k; return p; }

class SortedIndexBasedFilterOperator {
 public:
  static OpPtr create(const PredicateEvaluator& ev, const PbColumnView& c, int numDocs) {
    OpPtr op = mk(OP_SORTED);
    auto start = [&](int id) { return c.sorted_pairs[2 * id]; };
    auto end = [&](int id) { return c.sorted_pairs[2 * id + 1]; };
    if (ev.isRange()) { op->docIdRanges = {start(ev.startDictId), end(ev.endDictId - 1)}; return op; }
    std::vector<int32_t> r;
    int32_t ls = start(ev.dictIds[0]), le = end(ev.dictIds[0]);
    for (size_t i = 1; i < ev.dictIds.size(); i++) {
      int32_t s = start(ev.dictIds[i]), e = end(ev.dictIds[i]);
      if (s == le + 1) le = e; else { r.push_back(ls); r.push_back(le); ls = s; le = e; }
    }
    r.push_back(ls); r.push_back(le);
    if (ev.exclusive) {   // invert over [0, numDocs)
      std::vector<int32_t> inv;
      if (r[0] > 0) { inv.push_back(0); inv.push_back(r[0] - 1); }
      for (size_t i = 0; i + 2 < r.size(); i += 2) { inv.push_back(r[i + 1] + 1); inv.push_back(r[i + 2] - 1); }
      if (r[r.size() - 1] < numDocs - 1) { inv.push_back(r[r.size() - 1] + 1); inv.push_back(numDocs - 1); }
      r.swap(inv);
    }
    op->docIdRanges = r;
    return op;
  }
};

class FilterOperatorUtils {
 public:
  // getLeafFilterOperator: sorted > inverted > scan (RANGE never uses the inverted index); range-index/text/json/H3 are out of scope
  static OpPtr getLeafFilterOperator(const PredicateEvaluator& ev, const PbColumnView& c, int numDocs, bool skipInverted) {
    if (ev.alwaysFalse) return mk(OP_EMPTY);
    if (ev.alwaysTrue) return mk(OP_MATCH_ALL);
    if (c.is_sorted && c.has_dict) return SortedIndexBasedFilterOperator::create(ev, c, numDocs);
    OpPtr op;
    if (!ev.isRange() && c.has_inverted && c.has_dict && !skipInverted) op = mk(OP_INVERTED); else op = mk(OP_SCAN);
    op->ev = ev;
    return op;
  }
  static int getPriority(const FilterOperator& f) {   // PrioritizedFilterOperator constants
    switch (f.kind) {
      case OP_SORTED: return 0;
      case OP_INVERTED: return 100;
      case OP_BITMAP: return 100;     // BitmapBasedFilterOperator
      case OP_AND: return 300;
      case OP_OR: return 400;
      case OP_NOT: return getPriority(*f.children[0]);
      case OP_SCAN: return 500;
      default: return 10000;
    }
  }
  static OpPtr getAndFilterOperator(std::vector<OpPtr> ops) {
    std::vector<OpPtr> kids;
    for (auto& o : ops) if (o->kind == OP_EMPTY) return mk(OP_EMPTY);
    for (auto& o : ops) if (o->kind != OP_MATCH_ALL) kids.push_back(std::move(o));
    if (kids.empty()) return mk(OP_MATCH_ALL);
    if (kids.size() == 1) return std::move(kids[0]);
    std::stable_sort(kids.begin(), kids.end(), [](const OpPtr& a, const OpPtr& b) { return getPriority(*a) < getPriority(*b); });
    OpPtr op = mk(OP_AND); op->children = std::move(kids); return op;
  }
  static OpPtr getOrFilterOperator(std::vector<OpPtr> ops) {
    std::vector<OpPtr> kids;
    for (auto& o : ops) if (o->kind == OP_MATCH_ALL) return mk(OP_MATCH_ALL);
    for (auto& o : ops) if (o->kind != OP_EMPTY) kids.push_back(std::move(o));
    if (kids.empty()) return mk(OP_EMPTY);
    if (kids.size() == 1) return std::move(kids[0]);
    OpPtr op = mk(OP_OR); op->children = std::move(kids); return op;
  }
  static OpPtr getNotFilterOperator(OpPtr o) {
    if (o->kind == OP_MATCH_ALL) return mk(OP_EMPTY);
    if (o->kind == OP_EMPTY) return mk(OP_MATCH_ALL);
    OpPtr op = mk(OP_NOT); op->children.push_back(std::move(o)); return op;
  }
};

static int findColumn(const PbSegmentView& s, const char* name) {
  for (size_t i = 0; i < s.cols.size(); i++) if (s.cols[i].name == name) return (int)i;
  return -1;
}

class FilterPlanNode {
 public:
  // constructPhysicalOperator over the postfix FilterContext
  static OpPtr run(const PbSegmentView& seg, const pbh_query_context& q) { return run(seg, q, q.num_filter_nodes, q.filter_nodes, q.predicates); }
  // (the FILTER clause of a filtered aggregation is planned on its own: AggregationFunctionUtils.java:343-344)
  // IS NULL / IS NOT NULL leaf: a BitmapBasedFilterOperator over the column's null-value vector (exclusive for IS NOT NULL);
  // without a vector IS NULL is empty and IS NOT NULL matches all (FilterPlanNode.java:294-307)
  static OpPtr nullVectorOp(const PbSegmentView& seg, int ci, bool notNull) {
    const PbColumnView& nc = seg.cols[ci];
    if (!nc.null_vector) return mk(notNull ? OP_MATCH_ALL : OP_EMPTY);
    OpPtr op = mk(OP_BITMAP);
    op->bitmap = nc.null_vector; op->bitmapLen = nc.null_vector_len; op->bitmapExclusive = notNull; op->bitmapColumn = ci;
    return op;
  }
  static OpPtr leafOp(const PbSegmentView& seg, const pbh_query_context& q, const pbh_predicate& p, int ci) {
    if (p.type == PBH_IS_NULL || p.type == PBH_IS_NOT_NULL) return nullVectorOp(seg, ci, p.type == PBH_IS_NOT_NULL);
    bool skipInv = false;
    for (int k = 0; k < q.num_skip_inverted; k++) if (seg.cols[ci].name == q.skip_inverted_columns[k]) skipInv = true;
    PredicateEvaluator ev = PredicateEvaluatorProvider::getPredicateEvaluator(p, seg.cols[ci], ci);
    return FilterOperatorUtils::getLeafFilterOperator(ev, seg.cols[ci], seg.num_docs, skipInv);
  }

  // ---- enableNullHandling: three-valued doc sets of the operator tree (BaseFilterOperator.java:88-113 and the overrides in
  // BaseColumnFilterOperator.java:46-70, AndFilterOperator.java:52-88, OrFilterOperator.java:51-87, NotFilterOperator.java:
  // 52-63), folded into an ordinary operator tree:
  //   column leaf   trues = matches AND NOT nulls;  nulls = the null-value vector;  falses = NOT (trues OR nulls)
  //   IS [NOT] NULL, Empty, MatchAll: no nulls
  //   AND  trues = AND trues_i;  falses = NOT AND_i (trues_i OR nulls_i)        (nulls_i of DIRECT column-leaf children only:
  //   OR   trues = OR trues_i;   falses = NOT OR_i (trues_i OR nulls_i)          And / Or / Not do not override getNulls())
  //   NOT  trues = falses of the child;  falses = its trues
  struct Expr { int kind = 0; int pred = -1; std::vector<std::unique_ptr<Expr>> kids; };
  struct NullAware {
    const PbSegmentView& seg; const pbh_query_context& q; const pbh_predicate* preds;
    int col(const Expr& e) const {
      int ci = findColumn(seg, preds[e.pred].column);
      if (ci < 0) throw BadQuery{std::string("unknown column ") + preds[e.pred].column};
      return ci;
    }
    // the null-value vector of a column leaf that is a real column operator (not folded to Empty / MatchAll), else -1
    int nullsOf(const Expr& e) const {
      if (e.kind != PBH_PREDICATE) return -1;
      const pbh_predicate& p = preds[e.pred];
      if (p.type == PBH_IS_NULL || p.type == PBH_IS_NOT_NULL) return -1;
      const int ci = col(e);
      if (!seg.cols[ci].null_vector) return -1;
      OpPtr base = leafOp(seg, q, p, ci);
      return (base->kind == OP_EMPTY || base->kind == OP_MATCH_ALL) ? -1 : ci;
    }
    OpPtr truesOrNulls(const Expr& e) const {
      OpPtr t = trues(e);
      const int ci = nullsOf(e);
      if (ci < 0) return t;
      std::vector<OpPtr> kids;
      kids.push_back(std::move(t)); kids.push_back(nullVectorOp(seg, ci, false));
      return FilterOperatorUtils::getOrFilterOperator(std::move(kids));
    }
    OpPtr trues(const Expr& e) const {
      if (e.kind == PBH_NOT) return falses(*e.kids[0]);
      if (e.kind == PBH_PREDICATE) {
        const pbh_predicate& p = preds[e.pred];
        const int ci = col(e);
        OpPtr base = leafOp(seg, q, p, ci);
        if (p.type == PBH_IS_NULL || p.type == PBH_IS_NOT_NULL || base->kind == OP_EMPTY || !seg.cols[ci].null_vector) return base;
        // FilterOperatorUtils.java:78-88: an always-true predicate on a column with nulls becomes a BitmapBasedFilterOperator over
        // the flipped null bitmap (not a column operator: it reports no nulls of its own)
        if (base->kind == OP_MATCH_ALL) return nullVectorOp(seg, ci, true);
        std::vector<OpPtr> kids;                                     // excludeNulls: AND(matches, flip(nullBitmap))
        kids.push_back(std::move(base)); kids.push_back(nullVectorOp(seg, ci, true));
        return FilterOperatorUtils::getAndFilterOperator(std::move(kids));
      }
      std::vector<OpPtr> kids;
      for (auto& k : e.kids) kids.push_back(trues(*k));
      return e.kind == PBH_AND ? FilterOperatorUtils::getAndFilterOperator(std::move(kids)) : FilterOperatorUtils::getOrFilterOperator(std::move(kids));
    }
    OpPtr falses(const Expr& e) const {
      if (e.kind == PBH_NOT) return trues(*e.kids[0]);
      if (e.kind == PBH_PREDICATE) return FilterOperatorUtils::getNotFilterOperator(truesOrNulls(e));
      std::vector<OpPtr> kids;
      for (auto& k : e.kids) kids.push_back(truesOrNulls(*k));
      OpPtr inner = e.kind == PBH_AND ? FilterOperatorUtils::getAndFilterOperator(std::move(kids)) : FilterOperatorUtils::getOrFilterOperator(std::move(kids));
      return FilterOperatorUtils::getNotFilterOperator(std::move(inner));
    }
  };
  static OpPtr runNullHandling(const PbSegmentView& seg, const pbh_query_context& q, int num_filter_nodes, const pbh_filter_node* filter_nodes, const pbh_predicate* predicates) {
    std::vector<std::unique_ptr<Expr>> stack;
    for (int i = 0; i < num_filter_nodes; i++) {
      const pbh_filter_node& n = filter_nodes[i];
      std::unique_ptr<Expr> e(new Expr());
      e->kind = n.kind; e->pred = n.predicate;
      const int k = n.kind == PBH_PREDICATE ? 0 : n.kind == PBH_NOT ? 1 : n.num_children;
      if ((int)stack.size() < k || (n.kind != PBH_PREDICATE && k < 1)) throw BadQuery{"malformed filter"};
      for (size_t j = stack.size() - (size_t)k; j < stack.size(); j++) e->kids.push_back(std::move(stack[j]));
      stack.resize(stack.size() - (size_t)k);
      stack.push_back(std::move(e));
    }
    if (stack.size() != 1) throw BadQuery{"malformed filter"};
    NullAware na{seg, q, predicates};
    return na.trues(*stack[0]);
  }

  static OpPtr run(const PbSegmentView& seg, const pbh_query_context& q, int num_filter_nodes, const pbh_filter_node* filter_nodes, const pbh_predicate* predicates) {
    if (num_filter_nodes == 0) return mk(OP_MATCH_ALL);
    if (q.null_handling) return runNullHandling(seg, q, num_filter_nodes, filter_nodes, predicates);
    std::vector<OpPtr> stack;
    for (int i = 0; i < num_filter_nodes; i++) {
      const pbh_filter_node& n = filter_nodes[i];
      if (n.kind == PBH_PREDICATE) {
        const pbh_predicate& p = predicates[n.predicate];
        int ci = findColumn(seg, p.column);
        if (ci < 0) throw BadQuery{std::string("unknown column ") + p.column};
        stack.push_back(leafOp(seg, q, p, ci));
      } else if (n.kind == PBH_NOT) {
        if (stack.empty()) throw BadQuery{"malformed filter"};
        OpPtr c = std::move(stack.back()); stack.pop_back();
        stack.push_back(FilterOperatorUtils::getNotFilterOperator(std::move(c)));
      } else {
        if ((int)stack.size() < n.num_children || n.num_children < 1) throw BadQuery{"malformed filter"};
        std::vector<OpPtr> kids;
        for (size_t k = stack.size() - n.num_children; k < stack.size(); k++) kids.push_back(std::move(stack[k]));
        stack.resize(stack.size() - n.num_children);
        stack.push_back(n.kind == PBH_AND ? FilterOperatorUtils::getAndFilterOperator(std::move(kids)) : FilterOperatorUtils::getOrFilterOperator(std::move(kids)));
      }
    }
    if (stack.size() != 1) throw BadQuery{"malformed filter"};
    return std::move(stack[0]);
  }
};

// ---------------------------------------------------------------------------------------------- lowering to the C ABI
struct LoweredSegment {
  std::vector<pb_filter_node> nodes;
  std::vector<std::unique_ptr<std::vector<int32_t>>> idStore;
  std::vector<std::unique_ptr<std::vector<int64_t>>> rawStore;
};

static void emit(const FilterOperator& f, const PbSegmentView& seg, LoweredSegment& out) {
  pb_filter_node n;
  memset(&n, 0, sizeof n);
  switch (f.kind) {
    case OP_EMPTY: n.kind = PB_F_EMPTY; break;
    case OP_MATCH_ALL: n.kind = PB_F_MATCH_ALL; break;
    case OP_SORTED: {
      out.idStore.emplace_back(new std::vector<int32_t>(f.docIdRanges));
      n.kind = PB_F_SORTED; n.ids = out.idStore.back()->data(); n.num_ids = (int32_t)(f.docIdRanges.size() / 2);
      break;
    }
    case OP_INVERTED: {
      out.idStore.emplace_back(new std::vector<int32_t>(f.ev.dictIds));
      n.kind = PB_F_INVERTED; n.column = f.ev.column; n.exclusive = f.ev.exclusive; n.ids = out.idStore.back()->data(); n.num_ids = (int32_t)f.ev.dictIds.size();
      break;
    }
    case OP_BITMAP: n.kind = PB_F_BITMAP; n.column = f.bitmapColumn; n.exclusive = f.bitmapExclusive ? 1 : 0; break;   // no blob: the column's staged null-value vector
    case OP_SCAN: {
      const PredicateEvaluator& e = f.ev;
      n.column = e.column;
      if (e.dictionaryBased) {
        if (e.isRange()) { n.kind = PB_F_SCAN_DICT_RANGE; n.lo = e.startDictId; n.hi = e.endDictId; }
        else {
          out.idStore.emplace_back(new std::vector<int32_t>(e.dictIds));
          n.kind = PB_F_SCAN_DICT_SET; n.exclusive = e.exclusive; n.ids = out.idStore.back()->data(); n.num_ids = (int32_t)e.dictIds.size();
        }
      } else if (e.isRange()) {
        n.kind = PB_F_SCAN_RAW_RANGE; n.lo = e.ilo; n.hi = e.ihi; n.dlo = e.dlo; n.dhi = e.dhi; n.dlo_inclusive = e.dloIncl; n.dhi_inclusive = e.dhiIncl;
      } else {
        out.rawStore.emplace_back(new std::vector<int64_t>(e.rawValues));
        n.kind = PB_F_SCAN_RAW_SET; n.exclusive = e.exclusive; n.raw_values = out.rawStore.back()->data(); n.num_raw_values = (int32_t)e.rawValues.size();
      }
      break;
    }
    case OP_NOT: emit(*f.children[0], seg, out); n.kind = PB_F_NOT; n.num_children = 1; break;
    default:
      for (auto& c : f.children) emit(*c, seg, out);
      n.kind = f.kind == OP_AND ? PB_F_AND : PB_F_OR; n.num_children = (int32_t)f.children.size();
      break;
  }
  out.nodes.push_back(n);
}

static void explain(const FilterOperator& f, const PbSegmentView& seg, int depth, std::string& out) {
  static const char* names[] = {"FILTER_EMPTY", "FILTER_MATCH_ENTIRE_SEGMENT", "FILTER_SORTED_INDEX", "FILTER_INVERTED_INDEX", "FILTER_FULL_SCAN", "FILTER_AND", "FILTER_OR", "FILTER_NOT", "FILTER_BITMAP"};
  out.append((size_t)depth * 2, ' ');
  out += names[f.kind];
  if (f.kind == OP_SCAN || f.kind == OP_INVERTED) {
    char buf[160];
    static const char* pt[] = {"EQ", "NOT_EQ", "IN", "NOT_IN", "RANGE"};
    if (f.ev.dictionaryBased && f.ev.isRange()) snprintf(buf, sizeof buf, "(%s %s dictIds[%d,%d))", seg.cols[f.ev.column].name.c_str(), pt[f.ev.type], f.ev.startDictId, f.ev.endDictId);
    else snprintf(buf, sizeof buf, "(%s %s n=%zu)", seg.cols[f.ev.column].name.c_str(), pt[f.ev.type], f.ev.dictionaryBased ? f.ev.dictIds.size() : f.ev.rawValues.size());
    out += buf;
  }
  if (f.kind == OP_SORTED) { char buf[64]; snprintf(buf, sizeof buf, "(%zu docId ranges)", f.docIdRanges.size() / 2); out += buf; }
  out += "\n";
  for (auto& c : f.children) explain(*c, seg, depth + 1, out);
}

// ---------------------------------------------------------------------------------------------- B200PlanMaker
class B200PlanMaker {
 public:
  // the (segment, query) pairs this executor accepts; everything else keeps the stock CPU plan
  static void checkEligible(const PbSegmentView& seg, const pbh_query_context& q) {
    if (q.num_aggregations <= 0) throw BadQuery{"not an aggregation query"};
    for (int j = 0; j < q.num_group_by; j++) {
      int ci = findColumn(seg, q.group_by_columns[j]);
      if (ci < 0) throw BadQuery{std::string("unknown group-by column ") + q.group_by_columns[j]};
      if (!seg.cols[ci].has_dict && seg.cols[ci].type == PB_STRING) throw BadQuery{"raw STRING group-by key"};
      if (q.null_handling && seg.cols[ci].null_vector) throw BadQuery{"enableNullHandling with a nullable group-by column (null group keys stay on the CPU plan)"};
    }
    for (int a = 0; a < q.num_aggregations; a++) {
      const pb_aggregation_desc& ad = q.aggregations[a];
      if (ad.op < PB_AGG_COUNT || ad.op > PB_AGG_DISTINCTCOUNT) throw BadQuery{"aggregation function not offloaded"};
      if (ad.op == PB_AGG_COUNT && !(q.null_handling && ad.column)) continue;      // COUNT(col) = COUNT(*) unless nulls are handled
      int ci = ad.column ? findColumn(seg, ad.column) : -1;
      if (ci < 0) throw BadQuery{"unknown aggregation column"};
      if (ad.op == PB_AGG_COUNT) continue;
      if (ad.op == PB_AGG_DISTINCTCOUNT) { if (!seg.cols[ci].has_dict && seg.cols[ci].type == PB_STRING) throw BadQuery{"DISTINCTCOUNT on a raw STRING column"}; }
      else if (seg.cols[ci].type == PB_STRING) throw BadQuery{"numeric aggregation on STRING"};
    }
  }
};

// ---- enableNullHandling, aggregation side.  NullableSingleInputAggregationFunction.forEachNotNull / foldNotNull
// (CTR/query/aggregation/function/NullableSingleInputAggregationFunction.java:63-134) hand a function only the docs of a block
// whose input is not null.  On the device that is a FILTER clause: every aggregation over a column that has a null-value
// vector in some segment of the call gets the clause "<column> IS NOT NULL", ANDed with its own FILTER clause if it has one.
// Clauses are shared between functions with the same (own clause, column) pair, like the swim-lanes of filtered aggregations.
struct NullClause { int userClause; std::string column; };
struct NullClausePlan {
  std::vector<NullClause> clauses;
  std::vector<int32_t> of;            // per aggregation: index into clauses, -1 = none
};
static NullClausePlan planNullClauses(const std::vector<PbSegmentView>& views, const pbh_query_context& q) {
  NullClausePlan plan;
  plan.of.assign((size_t)q.num_aggregations, -1);
  for (int a = 0; a < q.num_aggregations; a++) {
    const int user = q.num_agg_filters > 0 ? q.agg_filter_of[a] : -1;
    std::string column;
    if (q.aggregations[a].column)
      for (const PbSegmentView& v : views) {
        const int ci = findColumn(v, q.aggregations[a].column);
        if (ci >= 0 && v.cols[ci].null_vector) { column = q.aggregations[a].column; break; }
      }
    if (user < 0 && column.empty()) continue;
    int hit = -1;
    for (size_t k = 0; k < plan.clauses.size(); k++) if (plan.clauses[k].userClause == user && plan.clauses[k].column == column) hit = (int)k;
    if (hit < 0) { plan.clauses.push_back({user, column}); hit = (int)plan.clauses.size() - 1; }
    plan.of[(size_t)a] = hit;
  }
  return plan;
}
static OpPtr nullClauseOp(const PbSegmentView& v, const pbh_query_context& q, const NullClause& c) {
  OpPtr op = c.userClause >= 0 ? FilterPlanNode::run(v, q, q.agg_filters[c.userClause].num_filter_nodes, q.agg_filters[c.userClause].filter_nodes, q.agg_filters[c.userClause].predicates)
                               : mk(OP_MATCH_ALL);
  if (c.column.empty()) return op;
  const int ci = findColumn(v, c.column.c_str());
  if (ci < 0) throw BadQuery{std::string("unknown aggregation column ") + c.column};
  std::vector<OpPtr> kids;
  kids.push_back(std::move(op)); kids.push_back(FilterPlanNode::nullVectorOp(v, ci, true));
  return FilterOperatorUtils::getAndFilterOperator(std::move(kids));
}

}  // namespace pinot_b200

using namespace pinot_b200;

extern "C" int pbh_is_eligible(pb_segment_group_handle g, const pbh_query_context* q) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q) return pbi_fail(PB_ERR_INVALID, "null query");
  try {
    for (auto s : segs) {
      PbSegmentView v;
      if ((rc = pbi_segment_view(s, &v))) return rc;
      B200PlanMaker::checkEligible(v, *q);
      FilterPlanNode::run(v, *q);
      for (int f = 0; f < q->num_agg_filters; f++) FilterPlanNode::run(v, *q, q->agg_filters[f].num_filter_nodes, q->agg_filters[f].filter_nodes, q->agg_filters[f].predicates);
    }
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
  return PB_OK;
}

extern "C" int pbh_execute(pb_segment_group_handle g, const pbh_query_context* q, uint32_t flags, pb_result_handle* out) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q || !out) return pbi_fail(PB_ERR_INVALID, "null argument");
  std::vector<LoweredSegment> lowered(segs.size());
  std::vector<pb_segment_query> sq(segs.size());
  if (q->num_agg_filters < 0 || (q->num_agg_filters > 0 && (!q->agg_filters || !q->agg_filter_of))) return pbi_fail(PB_ERR_INVALID, "bad FILTER clauses");
  // ---- plan cache, host side: the byte image of the unlowered query is the key; a hit replays the parked plan without
  // running FilterPlanNode / the predicate lowering for any segment ----
  std::string host_key;
  {
    auto put = [&](const void* p, size_t n) { host_key.append(static_cast<const char*>(p), n); };
    auto pod = [&](int64_t v) { put(&v, sizeof v); };
    auto str = [&](const char* c) { if (!c) { pod(-1); return; } const size_t n = strlen(c); pod((int64_t)n); put(c, n); };
    auto program = [&](int32_t n_nodes, const pbh_filter_node* nodes, const pbh_predicate* preds) {
      pod(n_nodes);
      for (int32_t i = 0; i < n_nodes; i++) {
        pod(nodes[i].kind); pod(nodes[i].num_children); pod(nodes[i].predicate);
        if (nodes[i].kind == PBH_PREDICATE) {
          const pbh_predicate& p = preds[nodes[i].predicate];
          pod(p.type); str(p.column); pod(p.num_values);
          for (int32_t k = 0; k < p.num_values; k++) str(p.values[k]);
          str(p.lower); str(p.upper); pod(p.lower_inclusive); pod(p.upper_inclusive);
        }
      }
    };
    pod(flags);
    program(q->num_filter_nodes, q->filter_nodes, q->predicates);
    pod(q->num_group_by);
    for (int32_t j = 0; j < q->num_group_by; j++) str(q->group_by_columns[j]);
    pod(q->num_aggregations);
    for (int32_t a = 0; a < q->num_aggregations; a++) { pod(q->aggregations[a].op); str(q->aggregations[a].column); }
    pod(q->num_groups_limit); pod(q->max_initial_result_holder_capacity);
    pod(q->num_skip_inverted);
    for (int32_t k = 0; k < q->num_skip_inverted; k++) str(q->skip_inverted_columns[k]);
    pod(q->num_agg_filters);
    for (int32_t f = 0; f < q->num_agg_filters; f++) program(q->agg_filters[f].num_filter_nodes, q->agg_filters[f].filter_nodes, q->agg_filters[f].predicates);
    if (q->num_agg_filters > 0) for (int32_t a = 0; a < q->num_aggregations; a++) pod(q->agg_filter_of[a]);
    pod(q->num_order_by); pod(q->trim_size); pod(q->trim_threshold); pod(q->null_handling);
    for (int32_t k = 0; k < q->num_order_by; k++) { pod(q->order_by[k].kind); pod(q->order_by[k].index); pod(q->order_by[k].descending); }
    pb_query_desc d0;
    memset(&d0, 0, sizeof d0);
    d0.flags = flags | (q->null_handling ? PB_Q_NULL_HANDLING : 0u);
    const int hit = pbi_plan_replay(g, host_key, &d0, out);
    if (hit != 0) return hit < 0 ? hit : PB_OK;
  }
  // views of all segments first: with enableNullHandling the clause list depends on which columns are nullable anywhere
  std::vector<PbSegmentView> views(segs.size());
  for (size_t i = 0; i < segs.size(); i++) if ((rc = pbi_segment_view(segs[i], &views[i]))) return rc;
  NullClausePlan nullPlan;
  int numClauses = q->num_agg_filters;
  const int32_t* clauseOf = q->agg_filter_of;
  std::vector<LoweredSegment> clauseLowered;
  std::vector<std::vector<const pb_filter_node*>> clausePtr(segs.size());
  std::vector<std::vector<int32_t>> clauseLen(segs.size());
  for (auto& sqi : sq) memset(&sqi, 0, sizeof sqi);
  try {
    if (q->null_handling) {
      nullPlan = planNullClauses(views, *q);
      numClauses = (int)nullPlan.clauses.size();
      clauseOf = nullPlan.of.data();
    }
    clauseLowered.resize(segs.size() * (size_t)numClauses);
    for (size_t i = 0; i < segs.size(); i++) {
      const PbSegmentView& v = views[i];
      B200PlanMaker::checkEligible(v, *q);
      OpPtr root = FilterPlanNode::run(v, *q);
      if (root->kind != OP_MATCH_ALL) emit(*root, v, lowered[i]);
      sq[i].filter = lowered[i].nodes.data();
      sq[i].num_filter_nodes = (int32_t)lowered[i].nodes.size();
      // FILTER clauses, each planned like a filter of its own for this segment
      for (int f = 0; f < numClauses; f++) {
        OpPtr sub;
        if (q->null_handling) sub = nullClauseOp(v, *q, nullPlan.clauses[(size_t)f]);
        else {
          const pbh_filter_program& fp = q->agg_filters[f];
          sub = FilterPlanNode::run(v, *q, fp.num_filter_nodes, fp.filter_nodes, fp.predicates);
        }
        LoweredSegment& ls = clauseLowered[i * (size_t)numClauses + f];
        if (sub->kind != OP_MATCH_ALL) emit(*sub, v, ls);
        clausePtr[i].push_back(ls.nodes.data());
        clauseLen[i].push_back((int32_t)ls.nodes.size());
      }
      sq[i].agg_filters = clausePtr[i].data();
      sq[i].agg_filter_nodes = clauseLen[i].data();
    }
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
  pb_query_desc d;
  memset(&d, 0, sizeof d);
  d.num_group_by = q->num_group_by; d.group_by_columns = q->group_by_columns;
  d.num_aggregations = q->num_aggregations; d.aggregations = q->aggregations;
  d.num_groups_limit = q->num_groups_limit > 0 ? q->num_groups_limit : 100000;
  d.max_initial_result_holder_capacity = q->max_initial_result_holder_capacity > 0 ? q->max_initial_result_holder_capacity : 10000;
  d.flags = flags | (q->null_handling ? PB_Q_NULL_HANDLING : 0u);
  d.num_agg_filters = numClauses; d.agg_filter_of = numClauses > 0 ? clauseOf : nullptr;
  d.num_order_by = q->num_order_by; d.order_by = q->order_by; d.trim_size = q->trim_size; d.trim_threshold = q->trim_threshold;
  pbi_set_pending_host_key(host_key);
  rc = pb_query_execute(g, sq.data(), &d, out);
  pbi_set_pending_host_key(std::string());
  return rc;
}

extern "C" int pbh_explain_filter(pb_segment_group_handle g, int32_t si, const pbh_query_context* q, char* buf, int32_t cap) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q || !buf || cap <= 0 || si < 0 || si >= (int)segs.size()) return pbi_fail(PB_ERR_INVALID, "bad argument");
  try {
    PbSegmentView v;
    if ((rc = pbi_segment_view(segs[si], &v))) return rc;
    OpPtr root = FilterPlanNode::run(v, *q);
    std::string s;
    explain(*root, v, 0, s);
    int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(buf, s.data(), (size_t)n); buf[n] = 0;
    return n;
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
}

extern "C" int pbh_explain_agg_filter(pb_segment_group_handle g, int32_t si, const pbh_query_context* q, int32_t clause, char* buf, int32_t cap) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q || !buf || cap <= 0 || si < 0 || si >= (int)segs.size() || clause < 0 || clause >= q->num_agg_filters) return pbi_fail(PB_ERR_INVALID, "bad argument");
  try {
    PbSegmentView v;
    if ((rc = pbi_segment_view(segs[si], &v))) return rc;
    const pbh_filter_program& fp = q->agg_filters[clause];
    OpPtr root = FilterPlanNode::run(v, *q, fp.num_filter_nodes, fp.filter_nodes, fp.predicates);
    std::string s;
    explain(*root, v, 0, s);
    int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(buf, s.data(), (size_t)n); buf[n] = 0;
    return n;
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
}

extern "C" int pbh_dump_lowered(pb_segment_group_handle g, int32_t si, const pbh_query_context* q, int32_t clause, char* buf, int32_t cap) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q || !buf || cap <= 0 || si < 0 || si >= (int)segs.size() || clause < -1) return pbi_fail(PB_ERR_INVALID, "bad argument");
  try {
    PbSegmentView v;
    if ((rc = pbi_segment_view(segs[si], &v))) return rc;
    OpPtr root;
    if (clause < 0) root = FilterPlanNode::run(v, *q);
    else if (q->null_handling) {
      // the clauses the device runs with enableNullHandling: (own FILTER clause, nullable input column) pairs, pbh_null_clause_plan
      std::vector<PbSegmentView> views(segs.size());
      for (size_t i = 0; i < segs.size(); i++) if ((rc = pbi_segment_view(segs[i], &views[i]))) return rc;
      NullClausePlan plan = planNullClauses(views, *q);
      if (clause >= (int)plan.clauses.size()) return pbi_fail(PB_ERR_INVALID, "bad clause index");
      root = nullClauseOp(v, *q, plan.clauses[(size_t)clause]);
    } else {
      if (clause >= q->num_agg_filters) return pbi_fail(PB_ERR_INVALID, "bad clause index");
      root = FilterPlanNode::run(v, *q, q->agg_filters[clause].num_filter_nodes, q->agg_filters[clause].filter_nodes, q->agg_filters[clause].predicates);
    }
    LoweredSegment ls;
    if (root->kind != OP_MATCH_ALL) emit(*root, v, ls);
    std::string s;
    char tmp[256];
    auto ids = [&](const int32_t* p, int n) { for (int i = 0; i < n; i++) { snprintf(tmp, sizeof tmp, i ? ",%d" : "%d", p[i]); s += tmp; } };
    for (const pb_filter_node& n : ls.nodes) {
      const char* col = (n.column >= 0 && n.column < (int)v.cols.size()) ? v.cols[n.column].name.c_str() : "";
      switch (n.kind) {
        case PB_F_AND: snprintf(tmp, sizeof tmp, "AND n=%d", n.num_children); s += tmp; break;
        case PB_F_OR: snprintf(tmp, sizeof tmp, "OR n=%d", n.num_children); s += tmp; break;
        case PB_F_NOT: s += "NOT"; break;
        case PB_F_MATCH_ALL: s += "MATCH_ALL"; break;
        case PB_F_EMPTY: s += "EMPTY"; break;
        case PB_F_SCAN_DICT_RANGE: snprintf(tmp, sizeof tmp, "SCAN_DICT_RANGE col=%s lo=%lld hi=%lld", col, (long long)n.lo, (long long)n.hi); s += tmp; break;
        case PB_F_SCAN_DICT_SET: case PB_F_INVERTED:
          snprintf(tmp, sizeof tmp, "%s col=%s excl=%d ids=", n.kind == PB_F_INVERTED ? "INVERTED" : "SCAN_DICT_SET", col, n.exclusive); s += tmp;
          ids(n.ids, n.num_ids);
          break;
        case PB_F_SCAN_RAW_RANGE:
          snprintf(tmp, sizeof tmp, "SCAN_RAW_RANGE col=%s ilo=%lld ihi=%lld dlo=%.17g dhi=%.17g dlo_incl=%d dhi_incl=%d", col, (long long)n.lo, (long long)n.hi,
                   n.dlo, n.dhi, n.dlo_inclusive, n.dhi_inclusive);
          s += tmp;
          break;
        case PB_F_SCAN_RAW_SET:
          snprintf(tmp, sizeof tmp, "SCAN_RAW_SET col=%s excl=%d vals=", col, n.exclusive); s += tmp;
          for (int i = 0; i < n.num_raw_values; i++) { snprintf(tmp, sizeof tmp, i ? ",%lld" : "%lld", (long long)n.raw_values[i]); s += tmp; }
          break;
        case PB_F_SORTED:
          s += "SORTED ranges=";
          for (int i = 0; i < n.num_ids; i++) { snprintf(tmp, sizeof tmp, i ? ",%d-%d" : "%d-%d", n.ids[2 * i], n.ids[2 * i + 1]); s += tmp; }
          break;
        case PB_F_BITMAP: snprintf(tmp, sizeof tmp, "BITMAP col=%s excl=%d %s", col, n.exclusive, n.blob ? "blob" : "null_value_vector"); s += tmp; break;
        default: s += "?"; break;
      }
      s += "\n";
    }
    int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    memcpy(buf, s.data(), (size_t)n); buf[n] = 0;
    return (int)s.size();
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
}

extern "C" int pbh_null_clause_plan(pb_segment_group_handle g, const pbh_query_context* q, int32_t* clause_of, int32_t cap) {
  std::vector<pb_segment_handle> segs;
  int rc = pbi_group_segments(g, &segs);
  if (rc) return rc;
  if (!q || (cap > 0 && !clause_of)) return pbi_fail(PB_ERR_INVALID, "bad argument");
  try {
    std::vector<PbSegmentView> views(segs.size());
    for (size_t i = 0; i < segs.size(); i++) if ((rc = pbi_segment_view(segs[i], &views[i]))) return rc;
    NullClausePlan plan;
    if (q->null_handling) plan = planNullClauses(views, *q);
    else { plan.of.assign((size_t)q->num_aggregations, -1); for (int a = 0; a < q->num_aggregations && q->num_agg_filters > 0; a++) plan.of[(size_t)a] = q->agg_filter_of[a]; }
    for (int a = 0; a < q->num_aggregations && a < cap; a++) clause_of[a] = plan.of[(size_t)a];
    return q->null_handling ? (int)plan.clauses.size() : q->num_agg_filters;
  } catch (const BadQuery& e) { return pbi_fail(PB_ERR_UNSUPPORTED, e.msg.c_str()); }
}
