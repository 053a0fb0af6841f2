/*
 * pinot_b200_host.h — C ABI of the host planning layer.
 *
 * In a Pinot server this layer is Java: B200PlanMaker extends InstancePlanMakerImplV2 and reuses the
 * reference's own PredicateEvaluatorProvider / FilterOperatorUtils to lower a QueryContext before calling
 * pinot_b200.h through JNI (java/ and INTEGRATION.md).  There is no JVM in this build environment, so the
 * same planning steps are written in C++ (pinot_b200/csrc/host/pb_host.cpp), class for class, and exposed
 * here so that tests and the benchmark can drive the device through exactly the path the plug-in takes:
 *
 *   QueryContext --(FilterPlanNode + PredicateEvaluator lowering, per segment)--> pb_segment_query[]
 *               --(B200PlanMaker eligibility)--> pb_query_execute
 *
 * CTR = pinot-core/src/main/java/org/apache/pinot/core
 */
#ifndef PINOT_B200_HOST_H
#define PINOT_B200_HOST_H

#include "pinot_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Predicate.Type (pinot-common/.../request/context/predicate/Predicate.java) */
enum { PBH_EQ = 0, PBH_NOT_EQ = 1, PBH_IN = 2, PBH_NOT_IN = 3, PBH_RANGE = 4,
       PBH_IS_NULL = 5, PBH_IS_NOT_NULL = 6 };   /* BitmapBasedFilterOperator over pb_column_desc.null_value_vector; Empty / MatchAll without one (CTR/plan/FilterPlanNode.java:294-307) */
/* FilterContext.Type */
enum { PBH_AND = 0, PBH_OR = 1, PBH_NOT = 2, PBH_PREDICATE = 3 };

/* A predicate as Pinot's QueryContext holds it: literals are strings
 * (EqPredicate/InPredicate/RangePredicate; RangePredicate.UNBOUNDED is a NULL bound here). */
typedef struct pbh_predicate {
  int32_t type;
  const char* column;
  int32_t num_values;
  const char* const* values;      /* EQ / NOT_EQ: 1; IN / NOT_IN: n */
  const char* lower;              /* RANGE; NULL = unbounded */
  const char* upper;
  int32_t lower_inclusive, upper_inclusive;
} pbh_predicate;

typedef struct pbh_filter_node {  /* postfix */
  int32_t kind;                   /* PBH_AND / PBH_OR (num_children operands) / PBH_NOT / PBH_PREDICATE */
  int32_t num_children;
  int32_t predicate;              /* index into predicates */
} pbh_filter_node;

/* One FILTER(WHERE ...) clause of a filtered aggregation (QueryContext.getFilteredAggregationFunctions()). */
typedef struct pbh_filter_program {
  int32_t num_filter_nodes;
  const pbh_filter_node* filter_nodes;
  const pbh_predicate* predicates;
} pbh_filter_program;

/* The part of QueryContext (CTR/query/request/context/QueryContext.java) this path reads. */
typedef struct pbh_query_context {
  int32_t num_filter_nodes;       /* 0 = no WHERE clause */
  const pbh_filter_node* filter_nodes;
  const pbh_predicate* predicates;
  int32_t num_group_by;
  const char* const* group_by_columns;
  int32_t num_aggregations;
  const pb_aggregation_desc* aggregations;
  int32_t num_groups_limit;
  int32_t max_initial_result_holder_capacity;
  int32_t num_skip_inverted;      /* query option skipIndexes: columns whose inverted index must not be used */
  const char* const* skip_inverted_columns;
  /* filtered aggregations: the distinct FILTER clauses (equal clauses share a swim-lane,
   * AggregationFunctionUtils.java:333-366) and, per aggregation, the index of its clause (-1 = none) */
  int32_t num_agg_filters;
  const pbh_filter_program* agg_filters;
  const int32_t* agg_filter_of;
  /* ORDER BY expressions that name a group-by column or an aggregation of the SELECT list, and the trim the combine layer
   * would apply (pb_query_desc.order_by / trim_size / trim_threshold; 0 = none) */
  int32_t num_order_by;
  const pb_order_by* order_by;
  int32_t trim_size;
  int32_t trim_threshold;
  /* query option enableNullHandling (QueryContext.isNullHandlingEnabled).  The filter is planned three-valued -- the trues of
   * the root, BaseFilterOperator.getTrues / getNulls / getFalses and the And / Or / Not / BaseColumnFilterOperator overrides --
   * and every aggregation whose input column has a null-value vector only sees its non-null docs
   * (NullableSingleInputAggregationFunction.java:63-134), which the device runs as the implicit FILTER clause
   * "<column> IS NOT NULL" (ANDed with the function's own clause); pb_result_long then carries the number of inputs every
   * function saw, 0 = SQL NULL (PB_Q_NULL_HANDLING).  Group-by columns with a null-value vector are declined. */
  int32_t null_handling;
} pbh_query_context;

/* B200PlanMaker.makeSegmentPlanNode eligibility (InstancePlanMakerImplV2.java:275-294 override):
 * returns PB_OK when every segment of the group can run on the device, PB_ERR_UNSUPPORTED (with the reason
 * in pb_last_error) when the plan maker must decline to the stock CPU plan. */
int pbh_is_eligible(pb_segment_group_handle g, const pbh_query_context* q);

/* Plan (FilterPlanNode.run per segment) + execute.  flags = PB_Q_* of pinot_b200.h.
 * Without PB_Q_COMBINE this is GroupByOperator/AggregationOperator.nextBlock() for each segment;
 * with it, the device-side equivalent of GroupByCombineOperator over the whole group. */
int pbh_execute(pb_segment_group_handle g, const pbh_query_context* q, uint32_t flags, pb_result_handle* out);

/* EXPLAIN-style dump of the lowered filter of one segment (Operator.toExplainString analogue); returns the
 * number of bytes written (excluding the terminator). */
int pbh_explain_filter(pb_segment_group_handle g, int32_t segment_index, const pbh_query_context* q, char* buf, int32_t cap);
/* Same for FILTER clause `clause` of a filtered aggregation (planned on its own, AggregationFunctionUtils.java:343-344). */
int pbh_explain_agg_filter(pb_segment_group_handle g, int32_t segment_index, const pbh_query_context* q, int32_t clause, char* buf, int32_t cap);

/* The lowered program itself — exactly the pb_filter_node list pbh_execute would hand to pb_query_execute for that segment
 * (clause = -1: the WHERE filter, >= 0: that FILTER clause) — as text, one postfix node per line with every dictId / docId /
 * raw value spelled out.  For tests and debugging of the lowering; returns the full length needed (excluding the terminator),
 * truncating to cap - 1 bytes. */
int pbh_dump_lowered(pb_segment_group_handle g, int32_t segment_index, const pbh_query_context* q, int32_t clause, char* buf, int32_t cap);

/* The FILTER clauses the device will run for `q` and, per aggregation, the index of its clause (-1 = none), written to
 * clause_of[0 .. min(cap, num_aggregations)).  Without enableNullHandling these are the query's own clauses; with it, one
 * clause per distinct (own clause, nullable input column) pair -- the function's clause ANDed with "<column> IS NOT NULL"
 * (see pbh_query_context.null_handling).  pbh_dump_lowered / clause indices refer to this list.  Returns the number of clauses. */
int pbh_null_clause_plan(pb_segment_group_handle g, const pbh_query_context* q, int32_t* clause_of, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
