/* Source only.  Replaces GroupByOperator / AggregationOperator (pinot-core/.../operator/query/GroupByOperator.java:101-140). */
package org.apache.pinot.b200;

import java.util.Collections;
import java.util.List;
import java.util.Map;
import org.apache.pinot.common.request.context.ExpressionContext;
import org.apache.pinot.common.request.context.FilterContext;
import org.apache.pinot.common.utils.DataSchema;
import org.apache.pinot.core.common.Operator;
import org.apache.pinot.core.operator.BaseOperator;
import org.apache.pinot.core.operator.ExecutionStatistics;
import org.apache.pinot.core.operator.blocks.results.GroupByResultsBlock;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.IndexSegment;

/**
 * nextBlock() is called once per segment by a combine worker (GroupByCombineOperator.java:110).  It stages the segment
 * on first use (B200SegmentCache keyed by the IndexSegment), calls Native.execute, and wraps the pinned result arrays:
 * DeviceGroupKeyGenerator.getGroupKeys() yields (groupId, decoded keys) from the device's decoded key columns;
 * DoubleGroupByResultHolder / ObjectGroupByResultHolder are filled from pb_result_double / pb_result_long so the stock
 * AggregationFunction.extractGroupByResult works (AggregationGroupByResult.java:54-56).
 * Extending BaseOperator keeps the interruption check and the trace scope (BaseOperator.java:38-53).
 */
public class B200GroupByOperator extends BaseOperator<GroupByResultsBlock> {
  private final IndexSegment _indexSegment;
  private final QueryContext _queryContext;
  private final B200FilterLowering.LoweredProgram _where;
  private final List<B200FilterLowering.LoweredProgram> _clauses;   // FILTER(WHERE ...) clauses, one program each
  private final Map<FilterContext, Integer> _clauseIndex;
  private final DataSchema _dataSchema;
  private long[] _stats = new long[5];

  public B200GroupByOperator(IndexSegment indexSegment, QueryContext queryContext, B200FilterLowering.LoweredProgram where,
      List<B200FilterLowering.LoweredProgram> clauses, Map<FilterContext, Integer> clauseIndex) {
    _indexSegment = indexSegment;
    _queryContext = queryContext;
    _where = where;
    _clauses = clauses;
    _clauseIndex = clauseIndex;
    // group-by columns before aggregation columns, as IndexedTable expects (GroupByOperator.java:74-96); the group-by
    // expressions are plain identifiers here (B200Eligibility), so their type is the column's own
    List<ExpressionContext> groupBy = queryContext.getGroupByExpressions();
    AggregationFunction[] functions = queryContext.getAggregationFunctions();
    String[] names = new String[groupBy.size() + functions.length];
    DataSchema.ColumnDataType[] types = new DataSchema.ColumnDataType[names.length];
    for (int i = 0; i < groupBy.size(); i++) {
      names[i] = groupBy.get(i).toString();
      types[i] = DataSchema.ColumnDataType.fromDataTypeSV(
          indexSegment.getDataSource(groupBy.get(i).getIdentifier()).getDataSourceMetadata().getDataType());
    }
    for (int i = 0; i < functions.length; i++) {
      names[groupBy.size() + i] = functions[i].getResultColumnName();
      types[groupBy.size() + i] = functions[i].getIntermediateResultColumnType();
    }
    _dataSchema = new DataSchema(names, types);
  }

  @Override
  protected GroupByResultsBlock getNextBlock() {
    // one-segment group: GroupByCombineOperator keeps merging per-segment blocks exactly as today.  (A combine-level
    // plug-in would hand all segments of the server to one call with PB_Q_COMBINE and skip that merge.)
    long group = B200SegmentCache.groupOf(_indexSegment);     // stages the segment on first use, cached until destroy()
    long result = B200Flatten.execute(group, _where, _clauses, _clauseIndex, _queryContext, /*flags=*/0);
    try {
      _stats = Native.resultStats(result, 0);
      return DeviceResults.toGroupByResultsBlock(result, _indexSegment, _queryContext, _dataSchema);
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
