/* Source only (no JDK in the build image). */
package org.apache.pinot.b200;

import java.nio.ByteBuffer;
import java.util.ArrayList;
import java.util.List;
import java.util.Map;
import java.util.concurrent.ConcurrentHashMap;
import org.apache.pinot.segment.local.indexsegment.immutable.ImmutableSegmentImpl;
import org.apache.pinot.segment.spi.ColumnMetadata;
import org.apache.pinot.segment.spi.IndexSegment;
import org.apache.pinot.segment.spi.index.StandardIndexes;
import org.apache.pinot.segment.spi.memory.PinotDataBuffer;
import org.apache.pinot.segment.spi.store.SegmentDirectory;
import org.apache.pinot.spi.data.FieldSpec.DataType;

/**
 * IndexSegment -> (staged segment handle, one-segment group handle).  Staging hands the library the segment's index buffers
 * exactly as they sit in columns.psf: SegmentDirectory.Reader.getIndexFor(column, type) (SPI/store/SegmentDirectory.java:179)
 * -> PinotDataBuffer.toDirectByteBuffer(0, size) (SPI/memory/PinotDataBuffer.java:654), a zero-copy view of the mmap.  The
 * library copies to HBM lazily, on first use by a query.  Entries are dropped from IndexSegment.destroy (the table data manager
 * releases a segment only after the last query on it has finished: BaseCombineOperator.java:86-90).
 */
final class B200SegmentCache {
  private B200SegmentCache() {
  }

  private static final class Entry {
    long _segment;
    long _group;
    List<String> _columns;
  }

  private static final Map<IndexSegment, Entry> ENTRIES = new ConcurrentHashMap<>();

  static long groupOf(IndexSegment segment) {
    return entry(segment)._group;
  }

  static List<String> stagedColumns(IndexSegment segment) {
    return entry(segment)._columns;
  }

  static void release(IndexSegment segment) {
    Entry e = ENTRIES.remove(segment);
    if (e != null) {
      Native.releaseGroup(e._group);
      Native.releaseSegment(e._segment);
    }
  }

  private static Entry entry(IndexSegment segment) {
    return ENTRIES.computeIfAbsent(segment, B200SegmentCache::stage);
  }

  private static Entry stage(IndexSegment indexSegment) {
    ImmutableSegmentImpl segment = (ImmutableSegmentImpl) indexSegment;          // eligibility admits immutable segments only
    SegmentDirectory.Reader reader = segment.getSegmentDirectory().createReader();
    List<String> columns = new ArrayList<>(segment.getSegmentMetadata().getColumnMetadataMap().keySet());
    int n = columns.size();
    int[] meta = new int[6 * n];
    ByteBuffer[] forward = new ByteBuffer[n];
    ByteBuffer[] dictionary = new ByteBuffer[n];
    ByteBuffer[] inverted = new ByteBuffer[n];
    ByteBuffer[] nullVectors = new ByteBuffer[n];      // IS NULL / IS NOT NULL leaves (FilterPlanNode: BitmapBasedFilterOperator)
    try {
      for (int i = 0; i < n; i++) {
        String column = columns.get(i);
        ColumnMetadata cm = segment.getSegmentMetadata().getColumnMetadataFor(column);
        DataType stored = cm.getDataType().getStoredType();
        meta[6 * i] = storedTypeCode(stored);
        meta[6 * i + 1] = cm.hasDictionary() ? 1 : 0;
        meta[6 * i + 2] = cm.isSorted() ? 1 : 0;
        meta[6 * i + 3] = cm.getCardinality();
        meta[6 * i + 4] = cm.getBitsPerElement();
        meta[6 * i + 5] = stored == DataType.STRING ? cm.getColumnMaxLength() : stored.size();
        forward[i] = view(reader.getIndexFor(column, StandardIndexes.forward()));
        if (cm.hasDictionary()) {
          dictionary[i] = view(reader.getIndexFor(column, StandardIndexes.dictionary()));
        }
        if (!cm.isSorted() && reader.hasIndexFor(column, StandardIndexes.inverted())) {
          inverted[i] = view(reader.getIndexFor(column, StandardIndexes.inverted()));
        }
        if (reader.hasIndexFor(column, StandardIndexes.nullValueVector())) {
          nullVectors[i] = view(reader.getIndexFor(column, StandardIndexes.nullValueVector()));
        }
      }
    } catch (java.io.IOException e) {
      throw new RuntimeException("cannot read the index buffers of " + segment.getSegmentName(), e);
    }
    Entry e = new Entry();
    e._columns = columns;
    e._segment = Native.stageSegment(segment.getSegmentName(), segment.getSegmentMetadata().getTotalDocs(),
        columns.toArray(new String[0]), meta, forward, dictionary, inverted, nullVectors);
    e._group = Native.createGroup(new long[]{e._segment});
    return e;
  }

  private static ByteBuffer view(PinotDataBuffer buffer) {
    return buffer.toDirectByteBuffer(0, (int) buffer.size());
  }

  /** PB_INT .. PB_STRING of include/pinot_b200.h */
  private static int storedTypeCode(DataType stored) {
    switch (stored) {
      case INT: return 0;
      case LONG: return 1;
      case FLOAT: return 2;
      case DOUBLE: return 3;
      case STRING: return 4;
      default: throw new B200Eligibility.NotEligibleException("stored type " + stored);
    }
  }
}
