/* Source only.  PlanNode { Operator run(); } (pinot-core/.../plan/PlanNode.java). */
package org.apache.pinot.b200;

import org.apache.pinot.core.operator.BaseOperator;
import org.apache.pinot.core.operator.blocks.results.BaseResultsBlock;
import org.apache.pinot.core.plan.FilterPlanNode;
import org.apache.pinot.core.plan.PlanNode;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.SegmentContext;

/**
 * run(): (1) FilterPlanNode.run() builds the reference's own filter operator tree (predicate evaluators, index
 * selection, AND re-ordering: FilterPlanNode.java:195-320, FilterOperatorUtils.java:74-252) while B200FilterOperatorUtils
 * records which evaluator each leaf came from; (2) B200FilterLowering walks that tree and emits the postfix pb_filter_node list (ScanBasedFilterOperator -> SCAN_DICT_RANGE / SCAN_DICT_SET / SCAN_RAW_*,
 * InvertedIndexFilterOperator -> INVERTED with getMatchingDictIds()/getNonMatchingDictIds(), SortedIndexBasedFilterOperator
 * -> SORTED docId ranges, BitmapBasedFilterOperator -> BITMAP); (3) returns a B200GroupByOperator / B200AggregationOperator.
 *
 * Filtered aggregations (QueryContext.getFilteredAggregationFunctions(), AGG(x) FILTER(WHERE ...)): every distinct
 * FilterContext gets a FilterPlanNode of its own, exactly as AggregationFunctionUtils.buildFilteredAggregationInfos does
 * (:343-344), and is lowered the same way into pb_segment_query.agg_filters[f]; pb_query_desc.agg_filter_of maps each
 * aggregation function to its clause.  The device evaluates the clauses per matching doc instead of running one
 * projection per swim-lane (DESIGN.md 4.5).
 */
public class B200AggregationPlanNode implements PlanNode {
  private final SegmentContext _segmentContext;
  private final QueryContext _queryContext;

  public B200AggregationPlanNode(SegmentContext segmentContext, QueryContext queryContext) {
    _segmentContext = segmentContext;
    _queryContext = queryContext;
  }

  @Override
  public BaseOperator<? extends BaseResultsBlock> run() {
    B200FilterLowering.LoweredProgram where = B200FilterLowering.lower(_segmentContext, _queryContext);
    // FILTER clauses of filtered aggregations: one program per distinct FilterContext, each planned with its own
    // FilterPlanNode(_segmentContext, _queryContext, filter)
    java.util.List<B200FilterLowering.LoweredProgram> clauses = new java.util.ArrayList<>();
    java.util.Map<org.apache.pinot.common.request.context.FilterContext, Integer> clauseIndex = new java.util.HashMap<>();
    if (_queryContext.getFilteredAggregationFunctions() != null) {
      for (org.apache.commons.lang3.tuple.Pair<?, org.apache.pinot.common.request.context.FilterContext> pair
          : _queryContext.getFilteredAggregationFunctions()) {
        if (pair.getRight() != null && !clauseIndex.containsKey(pair.getRight())) {
          clauseIndex.put(pair.getRight(), clauses.size());
          clauses.add(B200FilterLowering.lower(_segmentContext, _queryContext, pair.getRight()));
        }
      }
    }
    if (_queryContext.getGroupByExpressions() == null || _queryContext.getGroupByExpressions().isEmpty()) {
      return new B200AggregationOperator(_segmentContext.getIndexSegment(), _queryContext, where, clauses, clauseIndex);
    }
    return new B200GroupByOperator(_segmentContext.getIndexSegment(), _queryContext, where, clauses, clauseIndex);
  }
}
