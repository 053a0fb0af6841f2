/* Source only (no JDK in the build image). */
package org.apache.pinot.b200;

import java.util.ArrayList;
import java.util.List;
import java.util.Map;
import org.apache.commons.lang3.tuple.Pair;
import org.apache.pinot.common.request.context.ExpressionContext;
import org.apache.pinot.common.request.context.FilterContext;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.request.context.QueryContext;

/**
 * The lowered programs of one call -> the flattened arrays of Native.execute (layout: jni/pinot_b200_jni.c): nodes grouped by
 * (segment, program), program 0 = the WHERE filter, 1 + f = FILTER clause f; dictId / docId lists and raw values concatenated
 * into one pool each, with every node's offset rebased.
 */
final class B200Flatten {
  private B200Flatten() {
  }

  /** AggregationFunctionType -> PB_AGG_* */
  private static int aggOp(AggregationFunction<?, ?> function) {
    switch (function.getType()) {
      case COUNT: return 0;
      case SUM: return 1;
      case MIN: return 2;
      case MAX: return 3;
      case AVG: return 4;
      case DISTINCTCOUNT: return 5;
      default: throw new B200Eligibility.NotEligibleException("aggregation " + function.getType());
    }
  }

  /** one-segment call: `where` and `clauses` are the programs of that segment */
  static long execute(long group, B200FilterLowering.LoweredProgram where, List<B200FilterLowering.LoweredProgram> clauses,
      Map<FilterContext, Integer> clauseIndex, QueryContext queryContext, int flags) {
    List<B200FilterLowering.LoweredProgram> programs = new ArrayList<>();
    programs.add(where);
    programs.addAll(clauses);
    int numNodes = 0;
    int numIds = 0;
    int numRaws = 0;
    for (B200FilterLowering.LoweredProgram p : programs) {
      numNodes += p.numNodes();
      for (int[] ids : p._idLists) {
        numIds += ids.length;
      }
      for (long[] raws : p._rawLists) {
        numRaws += raws.length;
      }
    }
    final int ni = B200FilterLowering.LoweredProgram.INTS_PER_NODE;
    int[] nodeInts = new int[ni * numNodes];
    long[] nodeLongs = new long[2 * numNodes];
    double[] nodeDoubles = new double[2 * numNodes];
    int[] idPool = new int[Math.max(numIds, 1)];
    long[] rawPool = new long[Math.max(numRaws, 1)];
    int node = 0;
    int idOffset = 0;
    int rawOffset = 0;
    for (int prog = 0; prog < programs.size(); prog++) {
      B200FilterLowering.LoweredProgram p = programs.get(prog);
      for (int k = 0; k < p.numNodes(); k++, node++) {
        int[] v = p._ints.get(k).clone();
        v[0] = 0;                // segment (one-segment group)
        v[1] = prog;
        int[] ids = p._idLists.get(k);
        long[] raws = p._rawLists.get(k);
        v[7] = idOffset;
        v[9] = rawOffset;
        System.arraycopy(v, 0, nodeInts, ni * node, ni);
        System.arraycopy(ids, 0, idPool, idOffset, ids.length);
        System.arraycopy(raws, 0, rawPool, rawOffset, raws.length);
        idOffset += ids.length;
        rawOffset += raws.length;
        nodeLongs[2 * node] = p._longs.get(k)[0];
        nodeLongs[2 * node + 1] = p._longs.get(k)[1];
        nodeDoubles[2 * node] = p._doubles.get(k)[0];
        nodeDoubles[2 * node + 1] = p._doubles.get(k)[1];
      }
    }
    List<ExpressionContext> groupByExpressions = queryContext.getGroupByExpressions();
    String[] groupBy = new String[groupByExpressions == null ? 0 : groupByExpressions.size()];
    for (int j = 0; j < groupBy.length; j++) {
      groupBy[j] = groupByExpressions.get(j).getIdentifier();
    }
    AggregationFunction[] functions = queryContext.getAggregationFunctions();
    int[] aggOps = new int[functions.length];
    String[] aggColumns = new String[functions.length];
    int[] aggFilterOf = new int[functions.length];
    List<Pair<AggregationFunction, FilterContext>> filtered = queryContext.getFilteredAggregationFunctions();
    for (int a = 0; a < functions.length; a++) {
      aggOps[a] = aggOp(functions[a]);
      List<?> inputs = functions[a].getInputExpressions();
      ExpressionContext input = inputs.isEmpty() ? null : (ExpressionContext) inputs.get(0);
      aggColumns[a] = input != null && input.getType() == ExpressionContext.Type.IDENTIFIER && !"*".equals(input.getIdentifier())
          ? input.getIdentifier() : null;                       // COUNT(*) has no column
      FilterContext clause = filtered == null ? null : filtered.get(a).getRight();
      aggFilterOf[a] = clause == null ? -1 : clauseIndex.get(clause);
    }
    return Native.execute(group, 1, clauses.size(), nodeInts, nodeLongs, nodeDoubles, idPool, rawPool, groupBy, aggOps, aggColumns,
        aggFilterOf, queryContext.getNumGroupsLimit(), queryContext.getMaxInitialResultHolderCapacity(), flags);
  }
}
