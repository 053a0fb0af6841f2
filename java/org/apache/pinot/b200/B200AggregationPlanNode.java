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
 * selection, AND re-ordering: FilterPlanNode.java:195-320, FilterOperatorUtils.java:74-252); (2) B200FilterLowering walks
 * that tree and emits the postfix pb_filter_node list (ScanBasedFilterOperator -> SCAN_DICT_RANGE / SCAN_DICT_SET / SCAN_RAW_*,
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
    FilterPlanNode filterPlanNode = new FilterPlanNode(_segmentContext, _queryContext);
    long[] lowered = B200FilterLowering.lower(filterPlanNode.run(), _segmentContext.getIndexSegment());
    // FILTER clauses of filtered aggregations: B200FilterLowering.lowerClauses plans each distinct FilterContext with
    // new FilterPlanNode(_segmentContext, _queryContext, filter).run() and appends the lowered programs
    long[][] clauses = B200FilterLowering.lowerClauses(_segmentContext, _queryContext);
    return new B200GroupByOperator(_segmentContext.getIndexSegment(), _queryContext, lowered, clauses);
  }
}
