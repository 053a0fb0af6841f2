/*
 * Source only (no JDK in the build image).  Drop-in for pinot.server.query.executor.plan.maker.class
 * (pinot-core/.../query/config/QueryExecutorConfig.java:31,50; instantiated by ServerQueryExecutorV1Impl.java:116-123).
 */
package org.apache.pinot.b200;

import org.apache.pinot.core.plan.PlanNode;
import org.apache.pinot.core.plan.maker.InstancePlanMakerImplV2;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.core.query.request.context.utils.QueryContextUtils;
import org.apache.pinot.segment.spi.SegmentContext;

/**
 * Overrides only makeSegmentPlanNode (InstancePlanMakerImplV2.java:275-294): eligible (segment, query) pairs get a
 * B200 plan node, everything else keeps the stock CPU plan, so CombinePlanNode (CombinePlanNode.java:72-78) is unchanged.
 */
public class B200PlanMaker extends InstancePlanMakerImplV2 {
  @Override
  public void init(org.apache.pinot.spi.env.PinotConfiguration queryExecutorConfig) {
    super.init(queryExecutorConfig);
    // record (leaf operator -> predicate evaluator, data source) while the stock factory builds filter operators
    org.apache.pinot.core.operator.filter.FilterOperatorUtils.setImplementation(new B200FilterOperatorUtils());
  }

  @Override
  public PlanNode makeSegmentPlanNode(SegmentContext segmentContext, QueryContext queryContext) {
    if (QueryContextUtils.isAggregationQuery(queryContext) && B200Eligibility.isEligible(segmentContext, queryContext)) {
      return new B200AggregationPlanNode(segmentContext, queryContext);
    }
    return super.makeSegmentPlanNode(segmentContext, queryContext);
  }
}
