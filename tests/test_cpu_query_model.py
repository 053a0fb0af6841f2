"""The SQL subset front-end (tooling) — CPU only."""
import pytest

from pinot_b200.query import AggOp, And, Not, Or, Predicate, PredicateType, parse_sql, postfix_of


def test_filter_clauses_and_null_predicates():
    q = parse_sql("SELECT d, SUM(x) FILTER(WHERE a > 5 AND (b IN (1, 2) OR NOT c = 'z')), COUNT(*) FILTER(WHERE a IS NOT NULL), "
                  "AVG(y), MAX(x) FILTER ( WHERE a > 5 AND (b IN (1, 2) OR NOT c = 'z') ) FROM t WHERE e BETWEEN 1 AND 9 GROUP BY d LIMIT 7")
    assert [a.op for a in q.aggregations] == [AggOp.SUM, AggOp.COUNT, AggOp.AVG, AggOp.MAX]
    filters, index = q.agg_filters()
    assert index == [0, 1, -1, 0] and len(filters) == 2                    # equal clauses share one swim-lane
    f0 = filters[0]
    assert isinstance(f0, And) and isinstance(f0.children[1], Or) and isinstance(f0.children[1].children[1], Not)
    assert filters[1] == Predicate(PredicateType.IS_NOT_NULL, "a")
    nodes, preds = postfix_of(f0)
    assert [k for k, _, _ in nodes] == [3, 3, 3, 2, 1, 0] and [p.column for p in preds] == ["a", "b", "c"]
    assert q.filter == Predicate(PredicateType.RANGE, "e", lower="1", upper="9", lower_inclusive=True, upper_inclusive=True)
    assert q.group_by == ["d"] and q.limit == 7


def test_query_options():
    q = parse_sql("SET numGroupsLimit = 5; SET skipIndexes='c1=inverted,c2=inverted'; SELECT COUNT(*) FROM t WHERE c1 = 3")
    assert q.num_groups_limit == 5 and q.skip_indexes == {"c1": ["inverted"], "c2": ["inverted"]}
    with pytest.raises(ValueError):
        parse_sql("SET filteredAggregationsSkipEmptyGroups = 'true'; SELECT COUNT(*) FILTER(WHERE a > 1) FROM t")
    with pytest.raises(ValueError):
        parse_sql("SELECT a FROM t")                                        # selection queries are not on this path
    with pytest.raises(ValueError):
        parse_sql("SELECT PERCENTILE(a) FROM t")
