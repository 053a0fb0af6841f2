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
