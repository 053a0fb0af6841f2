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
import org.apache.pinot.core.operator.filter.BitmapBasedFilterOperator;
import org.apache.pinot.core.operator.filter.EmptyFilterOperator;
import org.apache.pinot.core.operator.filter.InvertedIndexFilterOperator;
import org.apache.pinot.core.operator.filter.MatchAllFilterOperator;
import org.apache.pinot.core.operator.filter.NotFilterOperator;
import org.apache.pinot.core.operator.filter.OrFilterOperator;
import org.apache.pinot.core.operator.filter.RangeIndexBasedFilterOperator;
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
 *   BitmapBasedFilterOperator (IS NULL / IS NOT NULL)       -> PB_F_BITMAP without a blob: the column's staged null-value vector
 *   MatchAllFilterOperator / EmptyFilterOperator            -> PB_F_MATCH_ALL / PB_F_EMPTY
 *   And / Or / NotFilterOperator                            -> PB_F_AND / PB_F_OR / PB_F_NOT after their children
 *   RangeIndexBasedFilterOperator                           -> the same leaf as the scan operator of that predicate (the index is exact;
 *                                                            the device scans the column instead of reading it)
 * Anything else (text / JSON / H3 index operators, expression filters) makes the segment ineligible.
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
    LoweredProgram program = new LoweredProgram(B200SegmentCache.stagedColumns(segmentContext.getIndexSegment()));
    program._segment = segmentContext.getIndexSegment();
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
    } else if (op instanceof BitmapBasedFilterOperator) {
      // FilterPlanNode.java:294-307 builds these itself for IS NULL / IS NOT NULL from DataSource.getNullValueVector() and the
      // operator keeps (bitmap, exclusive) private: B200FilterOperatorUtils.nullVectorLeaf recovers which staged column's null
      // bitmap it holds.  The device reads the vector staged with that column (pb_column_desc.null_value_vector).
      B200FilterOperatorUtils.NullVectorLeaf leaf = B200FilterOperatorUtils.nullVectorLeaf((BitmapBasedFilterOperator) op, out.segment(), out.stagedColumns());
      if (leaf == null) {
        throw new B200Eligibility.NotEligibleException("bitmap filter that is not a null-value vector");
      }
      out.addNullVector(leaf._column, leaf._exclusive);
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
      } else if ((op instanceof ScanBasedFilterOperator || op instanceof RangeIndexBasedFilterOperator) && ev.isDictionaryBased()) {
        if (ev instanceof SortedDictionaryBasedRangePredicateEvaluator) {
          SortedDictionaryBasedRangePredicateEvaluator range = (SortedDictionaryBasedRangePredicateEvaluator) ev;
          out.addDictIdRange(column, range.getStartDictId(), range.getEndDictId());
        } else {
          out.addDictIdSet(Native.PB_F_SCAN_DICT_SET, column, exclusive, exclusive ? ev.getNonMatchingDictIds() : ev.getMatchingDictIds());
        }
      } else if (op instanceof ScanBasedFilterOperator || op instanceof RangeIndexBasedFilterOperator) {
        // a range index (RangeIndexBasedFilterOperator.java: BitSlicedRangeIndexReader.getMatchingDocIds) returns exactly the
        // docs the predicate matches: the device gets the same set by scanning the column, which it does faster than the CPU
        // walks the bit slices, so the index is simply not read (numEntriesScannedInFilter then reports the scan)
        out.addRawPredicate(column, ev);        // inclusive bounds / value set of the raw-value evaluators
      } else {
        throw new B200Eligibility.NotEligibleException("filter operator " + op.toExplainString());
      }
    }
  }

  /**
   * One postfix program in the flattened layout of Native.execute: per node 12 ints (segment and program are filled in by
   * B200Flatten; kind, column, numChildren, exclusive, numIds, idOffset, numRaw, rawOffset, dloInclusive, dhiInclusive),
   * 2 longs (lo, hi) and 2 doubles (dlo, dhi); dictId / docId lists and raw values live in per-program pools whose offsets
   * B200Flatten rebases when it concatenates the programs of a call.
   */
  static final class LoweredProgram {
    static final int INTS_PER_NODE = 12;
    final List<int[]> _ints = new ArrayList<>();
    final List<long[]> _longs = new ArrayList<>();
    final List<double[]> _doubles = new ArrayList<>();
    final List<int[]> _idLists = new ArrayList<>();       // one entry per node (empty array when the node has none)
    final List<long[]> _rawLists = new ArrayList<>();
    private final List<String> _columns;                   // column order of the staged segment (B200SegmentCache)
    org.apache.pinot.segment.spi.IndexSegment _segment;    // the segment this program was lowered for

    org.apache.pinot.segment.spi.IndexSegment segment() {
      return _segment;
    }

    LoweredProgram() {
      this(new ArrayList<>());
    }

    LoweredProgram(List<String> stagedColumns) {
      _columns = stagedColumns;
    }

    int numNodes() {
      return _ints.size();
    }

    private void add(int kind, int column, int numChildren, boolean exclusive, int[] ids, int numIds, long[] raws, long lo, long hi,
        double dlo, double dhi, boolean dloInclusive, boolean dhiInclusive) {
      int[] node = new int[INTS_PER_NODE];
      node[2] = kind;
      node[3] = column;
      node[4] = numChildren;
      node[5] = exclusive ? 1 : 0;
      node[6] = numIds;                 // idOffset (7) and rawOffset (9) are assigned when the pools are concatenated
      node[8] = raws.length;
      node[10] = dloInclusive ? 1 : 0;
      node[11] = dhiInclusive ? 1 : 0;
      _ints.add(node);
      _longs.add(new long[]{lo, hi});
      _doubles.add(new double[]{dlo, dhi});
      _idLists.add(ids);
      _rawLists.add(raws);
    }

    void addCombinator(int kind, int numChildren) {
      add(kind, -1, numChildren, false, new int[0], 0, new long[0], 0, 0, 0, 0, false, false);
    }

    /** docIdRangePairs: inclusive (start, end) pairs, ascending; pb_filter_node.num_ids counts PAIRS for PB_F_SORTED. */
    void addSorted(int column, int[] docIdRangePairs) {
      add(Native.PB_F_SORTED, column, 0, false, docIdRangePairs, docIdRangePairs.length / 2, new long[0], 0, 0, 0, 0, false, false);
    }

    void addDictIdSet(int kind, int column, boolean exclusive, int[] dictIds) {
      int[] sorted = dictIds.clone();
      java.util.Arrays.sort(sorted);      // the C ABI wants ascending dictIds
      add(kind, column, 0, exclusive, sorted, sorted.length, new long[0], 0, 0, 0, 0, false, false);
    }

    /** IS NULL (exclusive = false) / IS NOT NULL (true): PB_F_BITMAP with no blob = the column's own null-value vector. */
    void addNullVector(int column, boolean exclusive) {
      add(Native.PB_F_BITMAP, column, 0, exclusive, new int[0], 0, new long[0], 0, 0, 0, 0, false, false);
    }

    List<String> stagedColumns() {
      return _columns;
    }

    void addDictIdRange(int column, int startDictId, int endDictIdExclusive) {
      add(Native.PB_F_SCAN_DICT_RANGE, column, 0, false, new int[0], 0, new long[0], startDictId, endDictIdExclusive, 0, 0, false, false);
    }

    /**
     * Raw-value (no-dictionary) scan leaves.  RANGE evaluators expose their bounds (Int/Long/Float/DoubleRawValueBasedRange-
     * PredicateEvaluator.getLowerBound() / getUpperBound(), already adjusted to INCLUSIVE bounds for the integral types:
     * RangePredicateEvaluatorFactory.java:331-366); EQ / IN / NOT_EQ / NOT_IN evaluators expose their value sets.
     */
    void addRawPredicate(int column, PredicateEvaluator evaluator) {
      Predicate.Type type = evaluator.getPredicateType();
      org.apache.pinot.spi.data.FieldSpec.DataType dataType = evaluator.getDataType();
      boolean integral = dataType == org.apache.pinot.spi.data.FieldSpec.DataType.INT
          || dataType == org.apache.pinot.spi.data.FieldSpec.DataType.LONG;
      if (type == Predicate.Type.RANGE) {
        org.apache.pinot.common.request.context.predicate.RangePredicate range =
            (org.apache.pinot.common.request.context.predicate.RangePredicate) evaluator.getPredicate();
        if (integral) {
          long lo = B200Eligibility.inclusiveLowerBound(evaluator);
          long hi = B200Eligibility.inclusiveUpperBound(evaluator);
          add(Native.PB_F_SCAN_RAW_RANGE, column, 0, false, new int[0], 0, new long[0], lo, hi, 0, 0, true, true);
        } else {
          double lo = range.getLowerBound().equals(org.apache.pinot.common.request.context.predicate.RangePredicate.UNBOUNDED)
              ? Double.NEGATIVE_INFINITY : Double.parseDouble(range.getLowerBound());
          double hi = range.getUpperBound().equals(org.apache.pinot.common.request.context.predicate.RangePredicate.UNBOUNDED)
              ? Double.POSITIVE_INFINITY : Double.parseDouble(range.getUpperBound());
          add(Native.PB_F_SCAN_RAW_RANGE, column, 0, false, new int[0], 0, new long[0], 0, 0, lo, hi,
              range.isLowerInclusive() || Double.isInfinite(lo), range.isUpperInclusive() || Double.isInfinite(hi));
        }
        return;
      }
      boolean exclusive = type == Predicate.Type.NOT_EQ || type == Predicate.Type.NOT_IN;
      long[] values = B200Eligibility.rawValueSet(evaluator, integral);      // doubles as IEEE-754 bits
      add(Native.PB_F_SCAN_RAW_SET, column, 0, exclusive, new int[0], 0, values, 0, 0, 0, 0, false, false);
    }

    /** position of the data source's column in the staged segment */
    int columnIndex(org.apache.pinot.segment.spi.datasource.DataSource dataSource) {
      String name = dataSource.getDataSourceMetadata().getFieldSpec().getName();
      int index = _columns.indexOf(name);
      if (index < 0) {
        throw new B200Eligibility.NotEligibleException("column " + name + " is not staged");
      }
      return index;
    }
  }
}
