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
  private final long[] _loweredFilter;
  private final long[][] _loweredClauses;   // FILTER(WHERE ...) clauses of filtered aggregations, one program each
  private long _numDocsScanned;
  private long _numEntriesScannedPostFilter;

  public B200GroupByOperator(IndexSegment indexSegment, QueryContext queryContext, long[] loweredFilter,
      long[][] loweredClauses) {
    _loweredClauses = loweredClauses;
    _indexSegment = indexSegment;
    _queryContext = queryContext;
    _loweredFilter = loweredFilter;
  }

  @Override
  protected GroupByResultsBlock getNextBlock() {
    long seg = B200SegmentCache.stage(_indexSegment);
    long result = Native.execute(seg, _loweredFilter, _loweredClauses, _queryContext);
    try {
      _numDocsScanned = Native.statNumDocsScanned(result);
      _numEntriesScannedPostFilter = Native.statNumEntriesScannedPostFilter(result);
      return DeviceResults.toGroupByResultsBlock(result, _indexSegment, _queryContext);
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
    return new ExecutionStatistics(_numDocsScanned, 0, _numEntriesScannedPostFilter,
        _indexSegment.getSegmentMetadata().getTotalDocs());
  }
}
