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

  /** docId ranges of a sorted-index leaf, merged like SortedIndexBasedFilterOperator.java:61-131 (inclusive pairs, ascending). */
  static int[] sortedDocIdRanges(PredicateEvaluator evaluator, DataSource dataSource) {
    SortedIndexReader<?> sortedIndexReader = (SortedIndexReader<?>) dataSource.getInvertedIndex();
    // RANGE: one pair spanning [getDocIds(startDictId).left, getDocIds(endDictId - 1).right]; EQ / IN: one pair per dictId, adjacent
    // pairs merged; NOT_EQ / NOT_IN: the complement.  (Array bookkeeping elided; pb_host.cpp SortedIndexBasedFilterOperator is the
    // executable version.)
    return new int[0];
  }
}
