/* Source only.  The keyless twin of B200GroupByOperator: replaces AggregationOperator (CTR/operator/query/AggregationOperator.java:64-80). */
package org.apache.pinot.b200;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;
import java.util.Collections;
import java.util.List;
import java.util.Map;
import org.apache.pinot.common.request.context.FilterContext;
import org.apache.pinot.core.common.Operator;
import org.apache.pinot.core.operator.BaseOperator;
import org.apache.pinot.core.operator.ExecutionStatistics;
import org.apache.pinot.core.operator.blocks.results.AggregationResultsBlock;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.local.customobject.AvgPair;
import org.apache.pinot.segment.spi.IndexSegment;

/**
 * One row of intermediate results, in the types AggregationFunction.extractAggregationResult would have produced
 * (CountAggregationFunction: Long; Sum / Min / Max: Double; Avg: AvgPair; DistinctCount: the value set), handed to
 * AggregationResultsBlock(functions, results, queryContext) exactly as AggregationOperator.getNextBlock does (:79).
 */
public class B200AggregationOperator extends BaseOperator<AggregationResultsBlock> {
  private final IndexSegment _indexSegment;
  private final QueryContext _queryContext;
  private final B200FilterLowering.LoweredProgram _where;
  private final List<B200FilterLowering.LoweredProgram> _clauses;
  private final Map<FilterContext, Integer> _clauseIndex;
  private long[] _stats = new long[5];

  public B200AggregationOperator(IndexSegment indexSegment, QueryContext queryContext, B200FilterLowering.LoweredProgram where,
      List<B200FilterLowering.LoweredProgram> clauses, Map<FilterContext, Integer> clauseIndex) {
    _indexSegment = indexSegment;
    _queryContext = queryContext;
    _where = where;
    _clauses = clauses;
    _clauseIndex = clauseIndex;
  }

  @Override
  protected AggregationResultsBlock getNextBlock() {
    long group = B200SegmentCache.groupOf(_indexSegment);
    long result = B200Flatten.execute(group, _where, _clauses, _clauseIndex, _queryContext, /*flags=*/0);
    try {
      _stats = Native.resultStats(result, 0);
      AggregationFunction[] functions = _queryContext.getAggregationFunctions();
      List<Object> results = new ArrayList<>(functions.length);
      for (int a = 0; a < functions.length; a++) {
        ByteBuffer doubles = Native.resultDoubles(result, 0, a).order(ByteOrder.nativeOrder());
        ByteBuffer longs = Native.resultLongs(result, 0, a).order(ByteOrder.nativeOrder());
        switch (functions[a].getType()) {
          case COUNT:
            results.add(longs.getLong(0));
            break;
          case SUM: case MIN: case MAX:
            results.add(doubles.getDouble(0));
            break;
          case AVG:
            results.add(new AvgPair(doubles.getDouble(0), longs.getLong(0)));
            break;
          default:
            results.add(DistinctSets.valueSet(result, a, 0, _indexSegment, functions[a]));
        }
      }
      return new AggregationResultsBlock(functions, results, _queryContext);
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
    return "AGGREGATE_B200";
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
