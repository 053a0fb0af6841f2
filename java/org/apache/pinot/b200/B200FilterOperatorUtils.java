/*
 * Source only (no JDK in the build image).
 */
package org.apache.pinot.b200;

import java.util.IdentityHashMap;
import java.util.List;
import java.util.Map;
import org.apache.pinot.core.operator.filter.BaseFilterOperator;
import org.apache.pinot.core.operator.filter.FilterOperatorUtils;
import org.apache.pinot.core.operator.filter.predicate.PredicateEvaluator;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.spi.datasource.DataSource;

/**
 * Pinot's leaf filter operators keep their PredicateEvaluator private, but the factory that creates them is pluggable:
 * FilterOperatorUtils.setImplementation (CTR/operator/filter/FilterOperatorUtils.java:41-70).  This implementation delegates
 * every decision to the stock DefaultImplementation (index selection :74-150, AND/OR simplification and priority order
 * :152-252) and only REMEMBERS which evaluator and data source each leaf operator was built from, per thread, so that
 * B200FilterLowering can turn the finished operator tree into pb_filter_node programs without re-deriving anything.
 * Installed once by B200PlanMaker.init().
 */
public final class B200FilterOperatorUtils implements FilterOperatorUtils.Implementation {
  static final class Leaf {
    final PredicateEvaluator _evaluator;
    final DataSource _dataSource;

    Leaf(PredicateEvaluator evaluator, DataSource dataSource) {
      _evaluator = evaluator;
      _dataSource = dataSource;
    }
  }

  /** A BitmapBasedFilterOperator that FilterPlanNode built from a column's null-value vector (IS NULL / IS NOT NULL). */
  static final class NullVectorLeaf {
    final int _column;          // index into the staged column list
    final boolean _exclusive;   // IS NOT NULL

    NullVectorLeaf(int column, boolean exclusive) {
      _column = column;
      _exclusive = exclusive;
    }
  }

  /**
   * FilterPlanNode.java:294-307 creates these operators itself (no FilterOperatorUtils call to intercept) and the operator
   * keeps its bitmap private, so the column is recovered by value: the operator's matching docs (getBitmaps().reduce()) equal
   * either some staged column's null bitmap (IS NULL) or its complement over [0, numDocs) (IS NOT NULL).  Null vectors are
   * small and this runs once per (segment, query).  Returns null when no column matches.
   */
  static NullVectorLeaf nullVectorLeaf(org.apache.pinot.core.operator.filter.BitmapBasedFilterOperator op,
      org.apache.pinot.segment.spi.IndexSegment segment, java.util.List<String> stagedColumns) {
    org.roaringbitmap.buffer.ImmutableRoaringBitmap trues = op.getBitmaps().reduce();
    int numDocs = segment.getSegmentMetadata().getTotalDocs();
    for (int i = 0; i < stagedColumns.size(); i++) {
      org.apache.pinot.segment.spi.index.reader.NullValueVectorReader reader =
          segment.getDataSource(stagedColumns.get(i)).getNullValueVector();
      if (reader == null) {
        continue;
      }
      org.roaringbitmap.buffer.ImmutableRoaringBitmap nulls = reader.getNullBitmap();
      if (nulls.equals(trues)) {
        return new NullVectorLeaf(i, false);
      }
      if (org.roaringbitmap.buffer.ImmutableRoaringBitmap.flip(nulls, 0L, numDocs).equals(trues)) {
        return new NullVectorLeaf(i, true);
      }
    }
    return null;
  }

  private static final ThreadLocal<Map<BaseFilterOperator, Leaf>> LEAVES = ThreadLocal.withInitial(IdentityHashMap::new);
  private final FilterOperatorUtils.Implementation _stock = new FilterOperatorUtils.DefaultImplementation();

  /** Leaves recorded on this thread since the last call (one FilterPlanNode.run() = one segment, one filter). */
  static Map<BaseFilterOperator, Leaf> takeLeaves() {
    Map<BaseFilterOperator, Leaf> leaves = LEAVES.get();
    LEAVES.set(new IdentityHashMap<>());
    return leaves;
  }

  @Override
  public BaseFilterOperator getLeafFilterOperator(QueryContext queryContext, PredicateEvaluator predicateEvaluator,
      DataSource dataSource, int numDocs) {
    BaseFilterOperator operator = _stock.getLeafFilterOperator(queryContext, predicateEvaluator, dataSource, numDocs);
    LEAVES.get().put(operator, new Leaf(predicateEvaluator, dataSource));
    return operator;
  }

  @Override
  public BaseFilterOperator getAndFilterOperator(QueryContext queryContext, List<BaseFilterOperator> filterOperators,
      int numDocs) {
    return _stock.getAndFilterOperator(queryContext, filterOperators, numDocs);
  }

  @Override
  public BaseFilterOperator getOrFilterOperator(QueryContext queryContext, List<BaseFilterOperator> filterOperators,
      int numDocs) {
    return _stock.getOrFilterOperator(queryContext, filterOperators, numDocs);
  }

  @Override
  public BaseFilterOperator getNotFilterOperator(QueryContext queryContext, BaseFilterOperator filterOperator,
      int numDocs) {
    return _stock.getNotFilterOperator(queryContext, filterOperator, numDocs);
  }
}
