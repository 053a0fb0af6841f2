/* Source only (no JDK in the build image). */
package org.apache.pinot.b200;

import it.unimi.dsi.fastutil.doubles.DoubleOpenHashSet;
import it.unimi.dsi.fastutil.floats.FloatOpenHashSet;
import it.unimi.dsi.fastutil.ints.IntOpenHashSet;
import it.unimi.dsi.fastutil.longs.LongOpenHashSet;
import it.unimi.dsi.fastutil.objects.ObjectOpenHashSet;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Set;
import org.apache.pinot.common.request.context.ExpressionContext;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.aggregation.groupby.GroupByResultHolder;
import org.apache.pinot.core.query.aggregation.groupby.ObjectGroupByResultHolder;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.segment.spi.index.reader.Dictionary;

/**
 * DISTINCTCOUNT intermediate results: the device hands back, per group, the ascending dictIds of the distinct values
 * (pb_result_distinct_offsets / _dict_ids); the function's intermediate result is the typed value set
 * (BaseDistinctAggregateAggregationFunction.java:760-806: IntOpenHashSet / LongOpenHashSet / FloatOpenHashSet /
 * DoubleOpenHashSet / ObjectOpenHashSet by stored type), which the combine layer merges by value.
 */
final class DistinctSets {
  private DistinctSets() {
  }

  static Set<?> valueSet(long result, int aggregation, int group, IndexSegment segment, AggregationFunction<?, ?> function) {
    ByteBuffer offsets = Native.resultDistinctOffsets(result, 0, aggregation).order(ByteOrder.nativeOrder());
    ByteBuffer dictIds = Native.resultDistinctDictIds(result, 0, aggregation).order(ByteOrder.nativeOrder());
    String column = ((ExpressionContext) function.getInputExpressions().get(0)).getIdentifier();
    Dictionary dictionary = segment.getDataSource(column).getDictionary();
    int from = (int) offsets.getLong(8 * group);
    int to = (int) offsets.getLong(8 * (group + 1));
    switch (dictionary.getValueType()) {
      case INT: {
        IntOpenHashSet set = new IntOpenHashSet(to - from);
        for (int k = from; k < to; k++) {
          set.add(dictionary.getIntValue(dictIds.getInt(4 * k)));
        }
        return set;
      }
      case LONG: {
        LongOpenHashSet set = new LongOpenHashSet(to - from);
        for (int k = from; k < to; k++) {
          set.add(dictionary.getLongValue(dictIds.getInt(4 * k)));
        }
        return set;
      }
      case FLOAT: {
        FloatOpenHashSet set = new FloatOpenHashSet(to - from);
        for (int k = from; k < to; k++) {
          set.add(dictionary.getFloatValue(dictIds.getInt(4 * k)));
        }
        return set;
      }
      case DOUBLE: {
        DoubleOpenHashSet set = new DoubleOpenHashSet(to - from);
        for (int k = from; k < to; k++) {
          set.add(dictionary.getDoubleValue(dictIds.getInt(4 * k)));
        }
        return set;
      }
      default: {
        ObjectOpenHashSet<String> set = new ObjectOpenHashSet<>(to - from);
        for (int k = from; k < to; k++) {
          set.add(dictionary.getStringValue(dictIds.getInt(4 * k)));
        }
        return set;
      }
    }
  }

  static GroupByResultHolder toHolder(long result, int aggregation, int numGroups, IndexSegment segment, AggregationFunction<?, ?> function) {
    ObjectGroupByResultHolder holder = new ObjectGroupByResultHolder(Math.max(numGroups, 1), Math.max(numGroups, 1));
    for (int g = 0; g < numGroups; g++) {
      holder.setValueForKey(g, valueSet(result, aggregation, g, segment, function));
    }
    return holder;
  }
}
