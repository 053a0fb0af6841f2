/* Source only.  Replaces GroupByOperator / AggregationOperator (pinot-core/.../operator/query/GroupByOperator.java:101-140). */
package org.apache.pinot.b200;

import java.util.Collections;
import java.util.List;
import org.apache.pinot.core.common.Operator;
import org.apache.pinot.core.operator.BaseOperator;
import org.apache.pinot.core.operator.ExecutionStatistics;
import org.apache.pinot.core.operator.blocks.results.GroupByResultsBlock;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.IndexSegment;

/**
 * nextBlock() is called once per segment by a combine worker (GroupByCombineOperator.java:110).  It stages the segment
 * on first use (B200SegmentCache keyed by segment name + CRC), calls Native.execute, and wraps the pinned result arrays:
 * DeviceGroupKeyGenerator.getGroupKeys() yields (groupId, decoded keys) from pb_result_group_dict_ids + the segment's own
 * Dictionary objects; DoubleGroupByResultHolder / ObjectGroupByResultHolder are filled from pb_result_double /
 * pb_result_long so the stock AggregationFunction.extractGroupByResult works (AggregationGroupByResult.java:54-56).
 * Extending BaseOperator keeps the interruption check and the trace scope (BaseOperator.java:38-53).
 */
public class B200GroupByOperator extends BaseOperator<GroupByResultsBlock> {
  private final IndexSegment _indexSegment;
  private final QueryContext _queryContext;
  private final B200FilterLowering.LoweredProgram _where;
  private final java.util.List<B200FilterLowering.LoweredProgram> _clauses;   // FILTER(WHERE ...) clauses, one program each
  private final java.util.Map<org.apache.pinot.common.request.context.FilterContext, Integer> _clauseIndex;
  private final org.apache.pinot.common.utils.DataSchema _dataSchema = null;   // built in the ctor exactly like GroupByOperator.java:65-98 (elided)
  private long[] _stats = new long[5];

  public B200GroupByOperator(IndexSegment indexSegment, QueryContext queryContext, B200FilterLowering.LoweredProgram where,
      java.util.List<B200FilterLowering.LoweredProgram> clauses,
      java.util.Map<org.apache.pinot.common.request.context.FilterContext, Integer> clauseIndex) {
    _indexSegment = indexSegment;
    _queryContext = queryContext;
    _where = where;
    _clauses = clauses;
    _clauseIndex = clauseIndex;
  }

  @Override
  protected GroupByResultsBlock getNextBlock() {
    // one-segment group: GroupByCombineOperator keeps merging per-segment blocks exactly as today.  (A combine-level
    // plug-in would hand all segments of the server to one call with PB_Q_COMBINE and skip that merge.)
    long group = B200SegmentCache.groupOf(_indexSegment);     // stages the segment on first use, cached until destroy()
    long result = B200Flatten.execute(group, _where, _clauses, _clauseIndex, _queryContext, /*flags=*/0);
    try {
      _stats = Native.resultStats(result, 0);
      // wraps the pinned arrays in GroupKeyGenerator / GroupByResultHolder implementations (AggregationGroupByResult.java:31-57)
      return DeviceResults.toGroupByResultsBlock(result, _indexSegment, _queryContext, _dataSchema);   // _dataSchema as in GroupByOperator.java:65-98
    } finally {
      Native.freeResult(result);
    }
  }

  @Override
  public List<Operator> getChildOperators() {
    return Collections.emptyList();
  }

  @Override
  public String toExplainString() {
    return "GROUP_BY_B200";
  }

  @Override
  public IndexSegment getIndexSegment() {
    return _indexSegment;
  }

  @Override
  public ExecutionStatistics getExecutionStatistics() {
    return new ExecutionStatistics(_stats[0], _stats[1], _stats[2], _stats[3]);
  }
}
