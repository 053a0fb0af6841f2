/*
 * Source only (no JDK in the build image).
 */
package org.apache.pinot.b200;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;
import java.util.Iterator;
import org.apache.pinot.common.utils.DataSchema;
import org.apache.pinot.core.operator.blocks.ValueBlock;
import org.apache.pinot.core.operator.blocks.results.GroupByResultsBlock;
import org.apache.pinot.core.query.aggregation.function.AggregationFunction;
import org.apache.pinot.core.query.aggregation.groupby.AggregationGroupByResult;
import org.apache.pinot.core.query.aggregation.groupby.DoubleGroupByResultHolder;
import org.apache.pinot.core.query.aggregation.groupby.GroupByResultHolder;
import org.apache.pinot.core.query.aggregation.groupby.GroupKeyGenerator;
import org.apache.pinot.core.query.aggregation.groupby.ObjectGroupByResultHolder;
import org.apache.pinot.core.query.request.context.QueryContext;
import org.apache.pinot.segment.local.customobject.AvgPair;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.spi.data.FieldSpec.DataType;

/**
 * The device result as the objects GroupByCombineOperator.processSegments consumes (GroupByCombineOperator.java:132-147):
 * a GroupKeyGenerator whose getGroupKeys() yields (row index, decoded keys) and one GroupByResultHolder per function,
 * so the stock AggregationFunction.extractGroupByResult works unchanged (AggregationGroupByResult.java:54-56).
 * Values are read from the pinned host arrays the finalize kernel wrote (native byte order).
 */
final class DeviceResults {
  private DeviceResults() {
  }

  static GroupByResultsBlock toGroupByResultsBlock(long result, IndexSegment segment, QueryContext queryContext, DataSchema dataSchema) {
    int numGroups = (int) Native.resultNumGroups(result, 0);
    AggregationFunction[] functions = queryContext.getAggregationFunctions();
    int numGroupBy = queryContext.getGroupByExpressions().size();
    // keys: decoded values, one column at a time
    Object[][] keys = new Object[numGroups][numGroupBy];
    for (int j = 0; j < numGroupBy; j++) {
      DataType storedType = segment.getDataSource(queryContext.getGroupByExpressions().get(j).getIdentifier()).getDataSourceMetadata()
          .getDataType().getStoredType();
      ByteBuffer values = Native.resultGroupKeyValues(result, 0, j).order(ByteOrder.nativeOrder());
      int width = numGroups == 0 ? 0 : values.capacity() / numGroups;
      for (int g = 0; g < numGroups; g++) {
        switch (storedType) {
          case INT: keys[g][j] = values.getInt(4 * g); break;
          case LONG: keys[g][j] = values.getLong(8 * g); break;
          case FLOAT: keys[g][j] = values.getFloat(4 * g); break;
          case DOUBLE: keys[g][j] = values.getDouble(8 * g); break;
          default: {               // STRING: fixed-width entry, zero padded (BaseImmutableDictionary.java:124-139)
            byte[] entry = new byte[width];
            values.position(width * g);
            values.get(entry);
            int len = width;
            while (len > 0 && entry[len - 1] == 0) {
              len--;
            }
            keys[g][j] = new String(entry, 0, len, StandardCharsets.UTF_8);
          }
        }
      }
    }
    GroupByResultHolder[] holders = new GroupByResultHolder[functions.length];
    for (int a = 0; a < functions.length; a++) {
      ByteBuffer doubles = Native.resultDoubles(result, 0, a).order(ByteOrder.nativeOrder());
      ByteBuffer longs = Native.resultLongs(result, 0, a).order(ByteOrder.nativeOrder());
      switch (functions[a].getType()) {
        case COUNT: case SUM: case MIN: case MAX: {
          DoubleGroupByResultHolder holder = new DoubleGroupByResultHolder(Math.max(numGroups, 1), Math.max(numGroups, 1), 0.0);
          for (int g = 0; g < numGroups; g++) {
            holder.setValueForKey(g, doubles.getDouble(8 * g));
          }
          holders[a] = holder;
          break;
        }
        case AVG: {
          ObjectGroupByResultHolder holder = new ObjectGroupByResultHolder(Math.max(numGroups, 1), Math.max(numGroups, 1));
          for (int g = 0; g < numGroups; g++) {
            holder.setValueForKey(g, new AvgPair(doubles.getDouble(8 * g), longs.getLong(8 * g)));
          }
          holders[a] = holder;
          break;
        }
        default: {                 // DISTINCTCOUNT: the value set, rebuilt from dictIds through the segment dictionary
          holders[a] = DistinctSets.toHolder(result, a, numGroups, segment, functions[a]);
        }
      }
    }
    AggregationGroupByResult groupByResult = new AggregationGroupByResult(new DeviceGroupKeyGenerator(keys), functions, holders);
    GroupByResultsBlock block = new GroupByResultsBlock(dataSchema, groupByResult, queryContext);
    block.setNumGroupsLimitReached(Native.resultStats(result, 0)[4] != 0);
    return block;
  }

  /** Row g of the device table is group id g. */
  private static final class DeviceGroupKeyGenerator implements GroupKeyGenerator {
    private final Object[][] _keys;

    DeviceGroupKeyGenerator(Object[][] keys) {
      _keys = keys;
    }

    @Override
    public int getGlobalGroupKeyUpperBound() {
      return _keys.length;
    }

    @Override
    public void generateKeysForBlock(ValueBlock valueBlock, int[] groupKeys) {
      throw new UnsupportedOperationException("keys were generated on the device");
    }

    @Override
    public void generateKeysForBlock(ValueBlock valueBlock, int[][] groupKeys) {
      throw new UnsupportedOperationException("keys were generated on the device");
    }

    @Override
    public int getCurrentGroupKeyUpperBound() {
      return _keys.length;
    }

    @Override
    public Iterator<GroupKey> getGroupKeys() {
      return new Iterator<GroupKey>() {
        private int _next;
        private final GroupKey _groupKey = new GroupKey();     // reused, as DictionaryBasedGroupKeyGenerator does

        @Override
        public boolean hasNext() {
          return _next < _keys.length;
        }

        @Override
        public GroupKey next() {
          _groupKey._groupId = _next;
          _groupKey._keys = _keys[_next++];
          return _groupKey;
        }
      };
    }

    @Override
    public int getNumKeys() {
      return _keys.length;
    }
  }
}
