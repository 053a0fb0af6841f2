/*
 * Source only (no JDK in the build image).  Executable twin: B200PlanMaker::checkEligible in pinot_b200/csrc/host/pb_host.cpp.
 */
package org.apache.pinot.b200;

import org.apache.pinot.common.request.context.ExpressionContext;
import org.apache.pinot.core.operator.filter.predicate.PredicateEvaluator;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.AggregationFunctionType;
import org.apache.pinot.segment.spi.ImmutableSegment;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.segment.spi.SegmentContext;
import org.apache.pinot.segment.spi.datasource.DataSource;
import org.apache.pinot.segment.spi.index.reader.SortedIndexReader;

/**
 * SURVEY.md 8b: which (segment, query) pairs run on the device.  Everything else keeps the stock CPU plan -- the plan maker
 * declines, it never falls back inside the operator.
 *   segment   immutable, no star-tree match for the query, no upsert validDocIds, no null handling
 *   group-by  identifiers of single-value columns; dictionary-encoded or raw fixed-width (no raw STRING/BYTES);
 *             dictId key space of at most 128 bits
 *   functions COUNT, SUM, MIN, MAX, AVG, DISTINCTCOUNT over single-value identifiers (DISTINCTCOUNT: dictionary column;
 *             no numeric function over STRING); at most 8 distinct FILTER(WHERE ...) clauses; not with
 *             filteredAggregationsSkipEmptyGroups
 *   filters   EQ, NOT_EQ, IN, NOT_IN, RANGE, IS [NOT] NULL on identifiers, combined with AND / OR / NOT; leaf operators
 *             chosen by the stock FilterOperatorUtils must be scan / inverted / sorted / match-all / empty
 *   forward   fixed-bit dictionary ids, sorted, or PASS_THROUGH raw chunks (compressed chunks decline)
 * The metadata fast paths of the reference stay in front: AggregationPlanNode still returns FastFilteredCountOperator /
 * NonScanBasedAggregationOperator when they apply (CTR/plan/AggregationPlanNode.java:95-140).
 */
final class B200Eligibility {
  static final class NotEligibleException extends RuntimeException {
    NotEligibleException(String reason) {
      super(reason);
    }
  }

  private B200Eligibility() {
  }

  static boolean isEligible(SegmentContext segmentContext, QueryContext queryContext) {
    IndexSegment segment = segmentContext.getIndexSegment();
    if (!(segment instanceof ImmutableSegment) || queryContext.isNullHandlingEnabled()) {
      return false;
    }
    if (queryContext.getGroupByExpressions() != null) {
      for (ExpressionContext expression : queryContext.getGroupByExpressions()) {
        if (expression.getType() != ExpressionContext.Type.IDENTIFIER || !isSingleValueFixedWidth(segment, expression.getIdentifier())) {
          return false;
        }
      }
    }
    for (AggregationFunction<?, ?> function : queryContext.getAggregationFunctions()) {
      AggregationFunctionType type = function.getType();
      if (type != AggregationFunctionType.COUNT && type != AggregationFunctionType.SUM && type != AggregationFunctionType.MIN
          && type != AggregationFunctionType.MAX && type != AggregationFunctionType.AVG
          && type != AggregationFunctionType.DISTINCTCOUNT) {
        return false;
      }
      for (Object input : function.getInputExpressions()) {
        ExpressionContext expression = (ExpressionContext) input;
        if (expression.getType() != ExpressionContext.Type.IDENTIFIER && type != AggregationFunctionType.COUNT) {
          return false;
        }
      }
    }
    try {
      // lowering doubles as the filter check: unsupported leaf operators throw NotEligibleException
      B200FilterLowering.lower(segmentContext, queryContext);
      return true;
    } catch (NotEligibleException e) {
      return false;
    }
  }

  private static boolean isSingleValueFixedWidth(IndexSegment segment, String column) {
    DataSource dataSource = segment.getDataSource(column);
    return dataSource.getDataSourceMetadata().isSingleValue()
        && (dataSource.getDictionary() != null || dataSource.getDataSourceMetadata().getDataType().getStoredType().isFixedWidth());
  }

  /**
   * docId ranges of a sorted-index leaf: inclusive (start, end) pairs in ascending order, adjacent ranges merged -- the list
   * SortedIndexBasedFilterOperator.getNextBlock builds (SortedIndexBasedFilterOperator.java:61-131).
   */
  static int[] sortedDocIdRanges(PredicateEvaluator evaluator, DataSource dataSource) {
    SortedIndexReader<?> sortedIndexReader = (SortedIndexReader<?>) dataSource.getInvertedIndex();
    java.util.List<int[]> ranges = new java.util.ArrayList<>();
    if (evaluator instanceof org.apache.pinot.core.operator.filter.predicate.RangePredicateEvaluatorFactory
        .SortedDictionaryBasedRangePredicateEvaluator) {
      org.apache.pinot.core.operator.filter.predicate.RangePredicateEvaluatorFactory.SortedDictionaryBasedRangePredicateEvaluator range =
          (org.apache.pinot.core.operator.filter.predicate.RangePredicateEvaluatorFactory.SortedDictionaryBasedRangePredicateEvaluator) evaluator;
      int start = range.getStartDictId();
      int end = range.getEndDictId();           // exclusive
      if (end > start) {
        ranges.add(new int[]{sortedIndexReader.getDocIds(start).getLeft(), sortedIndexReader.getDocIds(end - 1).getRight()});
      }
    } else {
      boolean exclusive = evaluator.getPredicateType() == org.apache.pinot.common.request.context.predicate.Predicate.Type.NOT_EQ
          || evaluator.getPredicateType() == org.apache.pinot.common.request.context.predicate.Predicate.Type.NOT_IN;
      int[] dictIds = exclusive ? evaluator.getNonMatchingDictIds() : evaluator.getMatchingDictIds();
      int[] sorted = dictIds.clone();
      java.util.Arrays.sort(sorted);
      java.util.List<int[]> hit = new java.util.ArrayList<>();
      for (int dictId : sorted) {
        org.apache.pinot.spi.utils.Pairs.IntPair docIds = sortedIndexReader.getDocIds(dictId);
        int[] last = hit.isEmpty() ? null : hit.get(hit.size() - 1);
        if (last != null && last[1] + 1 == docIds.getLeft()) {
          last[1] = docIds.getRight();           // adjacent dictIds cover adjacent docId ranges: merge
        } else {
          hit.add(new int[]{docIds.getLeft(), docIds.getRight()});
        }
      }
      if (!exclusive) {
        ranges = hit;
      } else {                                   // NOT_EQ / NOT_IN: the complement over [0, numDocs)
        int numDocs = dataSource.getDataSourceMetadata().getNumDocs();
        int next = 0;
        for (int[] r : hit) {
          if (r[0] > next) {
            ranges.add(new int[]{next, r[0] - 1});
          }
          next = r[1] + 1;
        }
        if (next < numDocs) {
          ranges.add(new int[]{next, numDocs - 1});
        }
      }
    }
    int[] flat = new int[2 * ranges.size()];
    for (int i = 0; i < ranges.size(); i++) {
      flat[2 * i] = ranges.get(i)[0];
      flat[2 * i + 1] = ranges.get(i)[1];
    }
    return flat;
  }

  /**
   * Inclusive bounds of a RANGE predicate on a raw INT / LONG column.  The raw-value evaluators that hold them
   * (RangePredicateEvaluatorFactory.Int/LongRawValueBasedRangePredicateEvaluator, :326-430) are private, so the bounds are
   * derived from the predicate's literals the same way the factory does: an exclusive bound moves by one, a fractional
   * bound rounds towards the inside of the interval, an unbounded side is the type's extreme.
   */
  static long inclusiveLowerBound(PredicateEvaluator evaluator) {
    org.apache.pinot.common.request.context.predicate.RangePredicate range =
        (org.apache.pinot.common.request.context.predicate.RangePredicate) evaluator.getPredicate();
    boolean isInt = evaluator.getDataType() == org.apache.pinot.spi.data.FieldSpec.DataType.INT;
    long min = isInt ? Integer.MIN_VALUE : Long.MIN_VALUE;
    if (range.getLowerBound().equals(org.apache.pinot.common.request.context.predicate.RangePredicate.UNBOUNDED)) {
      return min;
    }
    java.math.BigDecimal bound = new java.math.BigDecimal(range.getLowerBound());
    java.math.BigDecimal ceil = bound.setScale(0, java.math.RoundingMode.CEILING);
    java.math.BigInteger v = ceil.toBigInteger();
    if (ceil.compareTo(bound) == 0 && !range.isLowerInclusive()) {
      v = v.add(java.math.BigInteger.ONE);
    }
    return clamp(v, isInt);
  }

  static long inclusiveUpperBound(PredicateEvaluator evaluator) {
    org.apache.pinot.common.request.context.predicate.RangePredicate range =
        (org.apache.pinot.common.request.context.predicate.RangePredicate) evaluator.getPredicate();
    boolean isInt = evaluator.getDataType() == org.apache.pinot.spi.data.FieldSpec.DataType.INT;
    long max = isInt ? Integer.MAX_VALUE : Long.MAX_VALUE;
    if (range.getUpperBound().equals(org.apache.pinot.common.request.context.predicate.RangePredicate.UNBOUNDED)) {
      return max;
    }
    java.math.BigDecimal bound = new java.math.BigDecimal(range.getUpperBound());
    java.math.BigDecimal floor = bound.setScale(0, java.math.RoundingMode.FLOOR);
    java.math.BigInteger v = floor.toBigInteger();
    if (floor.compareTo(bound) == 0 && !range.isUpperInclusive()) {
      v = v.subtract(java.math.BigInteger.ONE);
    }
    return clamp(v, isInt);
  }

  private static long clamp(java.math.BigInteger v, boolean isInt) {
    java.math.BigInteger lo = java.math.BigInteger.valueOf(isInt ? Integer.MIN_VALUE : Long.MIN_VALUE);
    java.math.BigInteger hi = java.math.BigInteger.valueOf(isInt ? Integer.MAX_VALUE : Long.MAX_VALUE);
    return v.max(lo).min(hi).longValue();
  }

  /** the literal values of an EQ / NOT_EQ / IN / NOT_IN predicate on a raw column: longs, or IEEE-754 bits of the doubles */
  static long[] rawValueSet(PredicateEvaluator evaluator, boolean integral) {
    org.apache.pinot.common.request.context.predicate.Predicate predicate = evaluator.getPredicate();
    java.util.List<String> literals;
    switch (predicate.getType()) {
      case EQ:
        literals = java.util.Collections.singletonList(((org.apache.pinot.common.request.context.predicate.EqPredicate) predicate).getValue());
        break;
      case NOT_EQ:
        literals = java.util.Collections.singletonList(((org.apache.pinot.common.request.context.predicate.NotEqPredicate) predicate).getValue());
        break;
      case IN:
        literals = ((org.apache.pinot.common.request.context.predicate.InPredicate) predicate).getValues();
        break;
      default:
        literals = ((org.apache.pinot.common.request.context.predicate.NotInPredicate) predicate).getValues();
        break;
    }
    long[] out = new long[literals.size()];
    for (int i = 0; i < out.length; i++) {
      out[i] = integral ? new java.math.BigDecimal(literals.get(i)).longValueExact()
          : Double.doubleToLongBits(Double.parseDouble(literals.get(i)));
    }
    return out;
  }
}
