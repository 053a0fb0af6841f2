/*
 * Source only (no JDK in the build image).  Declarations match jni/pinot_b200_jni.c one to one.
 */
package org.apache.pinot.b200;

import java.nio.ByteBuffer;

final class Native {
  static {
    System.loadLibrary("pinot_b200_jni");     // links libpinot_b200.so
  }

  private Native() {
  }

  // pb_filter_node kinds (include/pinot_b200.h)
  static final int PB_F_AND = 0, PB_F_OR = 1, PB_F_NOT = 2, PB_F_MATCH_ALL = 3, PB_F_EMPTY = 4, PB_F_SCAN_DICT_RANGE = 5,
      PB_F_SCAN_DICT_SET = 6, PB_F_SCAN_RAW_RANGE = 7, PB_F_SCAN_RAW_SET = 8, PB_F_INVERTED = 9, PB_F_SORTED = 10, PB_F_BITMAP = 11;
  // pb_query_desc.flags
  static final int PB_Q_COMBINE = 1, PB_Q_GATHER_IN_PLACE = 16;

  /** meta: 6 ints per column (storedType, hasDictionary, isSorted, cardinality, bitsPerElement, dictEntryBytes); the buffers
   * are PinotDataBuffer.toDirectByteBuffer views of the mmap'd columns.psf (zero copy). */
  static native long stageSegment(String name, int numDocs, String[] columns, int[] meta, ByteBuffer[] forwardIndexes,
      ByteBuffer[] dictionaries, ByteBuffer[] invertedIndexes, ByteBuffer[] nullValueVectors);

  static native void releaseSegment(long segment);

  static native long createGroup(long[] segments);

  static native void releaseGroup(long group);

  /** Layout of the flattened programs: jni/pinot_b200_jni.c (12 ints, 2 longs, 2 doubles per node; nodes grouped by
   * (segment, program), program 0 = WHERE filter, 1 + f = FILTER clause f, postfix order inside a program). */
  static native long execute(long group, int numSegments, int numAggFilters, int[] nodeInts, long[] nodeLongs,
      double[] nodeDoubles, int[] idPool, long[] rawPool, String[] groupBy, int[] aggOps, String[] aggColumns,
      int[] aggFilterOf, int numGroupsLimit, int maxInitialResultHolderCapacity, int flags);

  static native long resultNumGroups(long result, int table);

  static native ByteBuffer resultDoubles(long result, int table, int aggregation);      // pinned host memory, valid until freeResult

  static native ByteBuffer resultLongs(long result, int table, int aggregation);

  static native ByteBuffer resultGroupDictIds(long result, int table, int groupByColumn);

  static native ByteBuffer resultGroupKeyValues(long result, int table, int groupByColumn);

  static native ByteBuffer resultDistinctOffsets(long result, int table, int aggregation);

  static native ByteBuffer resultDistinctDictIds(long result, int table, int aggregation);

  static native long[] resultStats(long result, int table);

  static native void freeResult(long result);
}
