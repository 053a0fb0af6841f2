/*
 * Source only (no JDK in the build image).  The executable specification of this class is its C++ twin,
 * pinot_b200/csrc/host/pb_host.cpp (emit()), which tests/ and bench.py drive through include/pinot_b200_host.h.
 */
package org.apache.pinot.b200;

import java.util.ArrayList;
import java.util.List;
import java.util.Map;
import org.apache.pinot.common.request.context.FilterContext;
import org.apache.pinot.common.request.context.predicate.Predicate;
import org.apache.pinot.core.common.Operator;
import org.apache.pinot.core.operator.filter.AndFilterOperator;
import org.apache.pinot.core.operator.filter.BaseFilterOperator;
import org.apache.pinot.core.operator.filter.EmptyFilterOperator;
import org.apache.pinot.core.operator.filter.InvertedIndexFilterOperator;
import org.apache.pinot.core.operator.filter.MatchAllFilterOperator;
import org.apache.pinot.core.operator.filter.NotFilterOperator;
import org.apache.pinot.core.operator.filter.OrFilterOperator;
import org.apache.pinot.core.operator.filter.ScanBasedFilterOperator;
import org.apache.pinot.core.operator.filter.SortedIndexBasedFilterOperator;
import org.apache.pinot.core.operator.filter.predicate.PredicateEvaluator;
import org.apache.pinot.core.operator.filter.predicate.RangePredicateEvaluatorFactory.SortedDictionaryBasedRangePredicateEvaluator;
import org.apache.pinot.core.plan.FilterPlanNode;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.SegmentContext;

/**
 * Operator tree (built by the stock FilterPlanNode) -> one postfix pb_filter_node program, in the flattened layout of
 * Native.execute (12 ints + 2 longs + 2 doubles per node, jni/pinot_b200_jni.c).  Leaf kinds:
 *   ScanBasedFilterOperator, dictionary column, RANGE      -> PB_F_SCAN_DICT_RANGE [getStartDictId(), getEndDictId())
 *   ScanBasedFilterOperator, dictionary column, EQ/IN/...  -> PB_F_SCAN_DICT_SET   getMatchingDictIds() or, for NOT_EQ / NOT_IN,
 *                                                            getNonMatchingDictIds() with exclusive = 1
 *   ScanBasedFilterOperator, raw column                    -> PB_F_SCAN_RAW_RANGE (getInclusiveLowerBound/UpperBound) / PB_F_SCAN_RAW_SET
 *   InvertedIndexFilterOperator                            -> PB_F_INVERTED  (same dictId lists; the device expands the bitmaps)
 *   SortedIndexBasedFilterOperator                         -> PB_F_SORTED    docId ranges from SortedIndexReader.getDocIds(dictId),
 *                                                            merged exactly as SortedIndexBasedFilterOperator.java:61-131 does
 *   MatchAllFilterOperator / EmptyFilterOperator            -> PB_F_MATCH_ALL / PB_F_EMPTY
 *   And / Or / NotFilterOperator                            -> PB_F_AND / PB_F_OR / PB_F_NOT after their children
 * Anything else (range / text / JSON / H3 index operators, expression filters) makes the segment ineligible.
 */
final class B200FilterLowering {
  private B200FilterLowering() {
  }

  /** The WHERE filter of the query for one segment. */
  static LoweredProgram lower(SegmentContext segmentContext, QueryContext queryContext) {
    return lower(segmentContext, queryContext, queryContext.getFilter());
  }

  /** One FILTER(WHERE ...) clause, planned on its own as AggregationFunctionUtils.buildFilteredAggregationInfos does (:343-344). */
  static LoweredProgram lower(SegmentContext segmentContext, QueryContext queryContext, FilterContext filter) {
    B200FilterOperatorUtils.takeLeaves();
    BaseFilterOperator root = new FilterPlanNode(segmentContext, queryContext, filter).run();
    Map<BaseFilterOperator, B200FilterOperatorUtils.Leaf> leaves = B200FilterOperatorUtils.takeLeaves();
    LoweredProgram program = new LoweredProgram();
    if (!(root instanceof MatchAllFilterOperator)) {       // an empty program means "matches all" to the device
      emit(root, leaves, program);
    }
    return program;
  }

  private static void emit(BaseFilterOperator op, Map<BaseFilterOperator, B200FilterOperatorUtils.Leaf> leaves,
      LoweredProgram out) {
    if (op instanceof AndFilterOperator || op instanceof OrFilterOperator) {
      List<Operator> children = op.getChildOperators();
      for (Operator child : children) {
        emit((BaseFilterOperator) child, leaves, out);
      }
      out.addCombinator(op instanceof AndFilterOperator ? Native.PB_F_AND : Native.PB_F_OR, children.size());
    } else if (op instanceof NotFilterOperator) {
      emit(((NotFilterOperator) op).getChildFilterOperator(), leaves, out);
      out.addCombinator(Native.PB_F_NOT, 1);
    } else if (op instanceof MatchAllFilterOperator) {
      out.addCombinator(Native.PB_F_MATCH_ALL, 0);
    } else if (op instanceof EmptyFilterOperator) {
      out.addCombinator(Native.PB_F_EMPTY, 0);
    } else {
      B200FilterOperatorUtils.Leaf leaf = leaves.get(op);
      if (leaf == null) {
        throw new B200Eligibility.NotEligibleException("filter operator " + op.toExplainString());
      }
      PredicateEvaluator ev = leaf._evaluator;
      int column = out.columnIndex(leaf._dataSource);
      boolean exclusive = ev.getPredicateType() == Predicate.Type.NOT_EQ || ev.getPredicateType() == Predicate.Type.NOT_IN;
      if (op instanceof SortedIndexBasedFilterOperator) {
        out.addSorted(column, B200Eligibility.sortedDocIdRanges(ev, leaf._dataSource));
      } else if (op instanceof InvertedIndexFilterOperator) {
        out.addDictIdSet(Native.PB_F_INVERTED, column, exclusive, exclusive ? ev.getNonMatchingDictIds() : ev.getMatchingDictIds());
      } else if (op instanceof ScanBasedFilterOperator && ev.isDictionaryBased()) {
        if (ev instanceof SortedDictionaryBasedRangePredicateEvaluator) {
          SortedDictionaryBasedRangePredicateEvaluator range = (SortedDictionaryBasedRangePredicateEvaluator) ev;
          out.addDictIdRange(column, range.getStartDictId(), range.getEndDictId());
        } else {
          out.addDictIdSet(Native.PB_F_SCAN_DICT_SET, column, exclusive, exclusive ? ev.getNonMatchingDictIds() : ev.getMatchingDictIds());
        }
      } else if (op instanceof ScanBasedFilterOperator) {
        out.addRawPredicate(column, ev);        // inclusive bounds / value set of the raw-value evaluators
      } else {
        throw new B200Eligibility.NotEligibleException("filter operator " + op.toExplainString());
      }
    }
  }

  /** Growable flattened program (see Native.execute). */
  static final class LoweredProgram {
    final List<int[]> _ints = new ArrayList<>();
    final List<long[]> _longs = new ArrayList<>();
    final List<double[]> _doubles = new ArrayList<>();
    final List<int[]> _idLists = new ArrayList<>();
    final List<long[]> _rawLists = new ArrayList<>();
    // addCombinator / addSorted / addDictIdSet / addDictIdRange / addRawPredicate / columnIndex fill one node each;
    // their bodies are array bookkeeping only.
    void addCombinator(int kind, int numChildren) { /* ... */ }
    void addSorted(int column, int[] docIdRangePairs) { /* ... */ }
    void addDictIdSet(int kind, int column, boolean exclusive, int[] dictIds) { /* ... */ }
    void addDictIdRange(int column, int startDictId, int endDictIdExclusive) { /* ... */ }
    void addRawPredicate(int column, PredicateEvaluator evaluator) { /* ... */ }
    int columnIndex(org.apache.pinot.segment.spi.datasource.DataSource dataSource) { return 0; /* position in the staged segment */ }
  }
}
