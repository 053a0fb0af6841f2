/*
 * pinot_b200_jni.c — thin JNI shim over include/pinot_b200.h (source only: there is no JDK / jni.h in the build
 * image; build on a Pinot server box with
 *     gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/pinot_b200_jni.c \
 *         -Lpinot_b200 -lpinot_b200 -o libpinot_b200_jni.so )
 *
 * Conventions: direct ByteBuffers come from PinotDataBuffer.toDirectByteBuffer(offset, size)
 * (SPI/memory/PinotDataBuffer.java:654) so GetDirectBufferAddress is a zero-copy view of the mmap; descriptors are
 * passed as flattened primitive arrays built by org.apache.pinot.b200.Native; every failure becomes a
 * RuntimeException carrying pb_last_error() (the combine layer wraps it with the segment name,
 * CTR/operator/combine/BaseCombineOperator.java:185-199).
 */
#ifdef PB_WITH_JNI
#include <jni.h>
#include <stdlib.h>
#include <string.h>
#include "pinot_b200.h"

static void throw_last(JNIEnv* env) {
  jclass c = (*env)->FindClass(env, "java/lang/RuntimeException");
  (*env)->ThrowNew(env, c, pb_last_error());
}

/* long stageSegment(String name, int numDocs, String[] colNames, int[] meta /\* 6 ints per column: storedType,
 * hasDictionary, isSorted, cardinality, bitsPerElement, dictEntryBytes *\/, ByteBuffer[] fwd, ByteBuffer[] dict,
 * ByteBuffer[] inv, ByteBuffer[] nullVectors) */
JNIEXPORT jlong JNICALL Java_org_apache_pinot_b200_Native_stageSegment(JNIEnv* env, jclass cls, jstring name, jint numDocs,
    jobjectArray colNames, jintArray meta, jobjectArray fwd, jobjectArray dict, jobjectArray inv, jobjectArray nullVectors) {
  jsize n = (*env)->GetArrayLength(env, colNames);
  pb_column_desc* cols = (pb_column_desc*)calloc((size_t)n, sizeof(pb_column_desc));
  jint* m = (*env)->GetIntArrayElements(env, meta, NULL);
  const char** names = (const char**)calloc((size_t)n, sizeof(char*));
  jstring* jnames = (jstring*)calloc((size_t)n, sizeof(jstring));
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, colNames, i);
    jnames[i] = s;
    names[i] = (*env)->GetStringUTFChars(env, s, NULL);
    cols[i].name = names[i];
    cols[i].stored_type = m[6 * i]; cols[i].has_dictionary = m[6 * i + 1]; cols[i].is_sorted = m[6 * i + 2];
    cols[i].cardinality = m[6 * i + 3]; cols[i].bits_per_element = m[6 * i + 4]; cols[i].dict_entry_bytes = m[6 * i + 5];
    jobject b = (*env)->GetObjectArrayElement(env, fwd, i);
    cols[i].forward_index = (*env)->GetDirectBufferAddress(env, b);
    cols[i].forward_index_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b);
    (*env)->DeleteLocalRef(env, b);      /* the ByteBuffer stays referenced by the array; a wide table must not fill the local-ref table */
    b = (*env)->GetObjectArrayElement(env, dict, i);
    if (b) { cols[i].dictionary = (*env)->GetDirectBufferAddress(env, b); cols[i].dictionary_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b); (*env)->DeleteLocalRef(env, b); }
    b = (*env)->GetObjectArrayElement(env, inv, i);
    if (b) { cols[i].inverted_index = (*env)->GetDirectBufferAddress(env, b); cols[i].inverted_index_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b); (*env)->DeleteLocalRef(env, b); }
    b = (*env)->GetObjectArrayElement(env, nullVectors, i);
    if (b) { cols[i].null_value_vector = (*env)->GetDirectBufferAddress(env, b); cols[i].null_value_vector_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b); (*env)->DeleteLocalRef(env, b); }
  }
  const char* sname = (*env)->GetStringUTFChars(env, name, NULL);
  pb_segment_desc d = { sname, numDocs, (int32_t)n, cols };
  pb_segment_handle h = NULL;
  int rc = pb_segment_stage(&d, 0, &h);       /* copies the names; the buffers stay owned by Pinot (mmap) */
  (*env)->ReleaseStringUTFChars(env, name, sname);
  for (jsize i = 0; i < n; i++) { (*env)->ReleaseStringUTFChars(env, jnames[i], names[i]); (*env)->DeleteLocalRef(env, jnames[i]); }
  (*env)->ReleaseIntArrayElements(env, meta, m, JNI_ABORT);
  free(jnames); free(names); free(cols);
  if (rc != PB_OK) { throw_last(env); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_org_apache_pinot_b200_Native_releaseSegment(JNIEnv* env, jclass cls, jlong h) {
  pb_segment_release((pb_segment_handle)(intptr_t)h);
}

JNIEXPORT jlong JNICALL Java_org_apache_pinot_b200_Native_createGroup(JNIEnv* env, jclass cls, jlongArray segs) {
  jsize n = (*env)->GetArrayLength(env, segs);
  jlong* p = (*env)->GetLongArrayElements(env, segs, NULL);
  pb_segment_handle* hs = (pb_segment_handle*)calloc((size_t)n, sizeof(*hs));
  for (jsize i = 0; i < n; i++) hs[i] = (pb_segment_handle)(intptr_t)p[i];
  pb_segment_group_handle g = NULL;
  int rc = pb_segment_group_create(hs, (int)n, &g);
  (*env)->ReleaseLongArrayElements(env, segs, p, JNI_ABORT);
  free(hs);
  if (rc != PB_OK) { throw_last(env); return 0; }
  return (jlong)(intptr_t)g;
}

/* long execute(long group, int numSegments, int numAggFilters,
 *              int[] nodeInts, long[] nodeLongs, double[] nodeDoubles, int[] idPool, long[] rawPool,
 *              String[] groupBy, int[] aggOps, String[] aggCols, int[] aggFilterOf,
 *              int numGroupsLimit, int maxInitCapacity, int flags)
 *
 * The lowered filter programs of ALL segments, flattened by B200FilterLowering (Native.java documents the same layout):
 *   nodeInts    12 ints per node: segment, program (0 = the WHERE filter, 1 + f = FILTER clause f), kind (PB_F_*), column,
 *               numChildren, exclusive, numIds, idOffset (into idPool; SORTED: 2 * numIds ints), numRaw, rawOffset (into
 *               rawPool), dloInclusive, dhiInclusive
 *   nodeLongs   2 per node: lo, hi          nodeDoubles  2 per node: dlo, dhi
 * Nodes arrive grouped by (segment, program), each program in postfix order — the order FilterPlanNode's operator tree is
 * walked in.  PB_F_BITMAP leaves (caller-supplied bitmaps) are not produced by the Java side. */
JNIEXPORT jlong JNICALL Java_org_apache_pinot_b200_Native_execute(JNIEnv* env, jclass cls, jlong group, jint numSegments,
    jint numAggFilters, jintArray nodeInts, jlongArray nodeLongs, jdoubleArray nodeDoubles, jintArray idPool, jlongArray rawPool,
    jobjectArray groupBy, jintArray aggOps, jobjectArray aggCols, jintArray aggFilterOf, jint numGroupsLimit, jint maxInitCapacity,
    jint flags) {
  enum { NI = 12 };
  const jsize nNodes = (*env)->GetArrayLength(env, nodeInts) / NI;
  jint* ni = (*env)->GetIntArrayElements(env, nodeInts, NULL);
  jlong* nl = (*env)->GetLongArrayElements(env, nodeLongs, NULL);
  jdouble* nd = (*env)->GetDoubleArrayElements(env, nodeDoubles, NULL);
  jint* ids = (*env)->GetIntArrayElements(env, idPool, NULL);
  jlong* raws = (*env)->GetLongArrayElements(env, rawPool, NULL);
  const int nProg = 1 + numAggFilters;
  pb_filter_node* nodes = (pb_filter_node*)calloc((size_t)(nNodes > 0 ? nNodes : 1), sizeof(pb_filter_node));
  pb_segment_query* sq = (pb_segment_query*)calloc((size_t)numSegments, sizeof(pb_segment_query));
  /* per (segment, program): first node and node count */
  const pb_filter_node** progFirst = (const pb_filter_node**)calloc((size_t)numSegments * nProg, sizeof(void*));
  int32_t* progLen = (int32_t*)calloc((size_t)numSegments * nProg, sizeof(int32_t));
  for (jsize i = 0; i < nNodes; i++) {
    const jint* v = ni + (size_t)i * NI;
    pb_filter_node* n = &nodes[i];
    n->kind = v[2]; n->column = v[3]; n->num_children = v[4]; n->exclusive = v[5];
    n->num_ids = v[6]; n->ids = v[6] > 0 ? (const int32_t*)(ids + v[7]) : NULL;
    n->num_raw_values = v[8]; n->raw_values = v[8] > 0 ? (const int64_t*)(raws + v[9]) : NULL;
    n->dlo_inclusive = v[10]; n->dhi_inclusive = v[11];
    n->lo = nl[2 * i]; n->hi = nl[2 * i + 1]; n->dlo = nd[2 * i]; n->dhi = nd[2 * i + 1];
    const size_t slot = (size_t)v[0] * nProg + (size_t)v[1];
    if (progLen[slot] == 0) progFirst[slot] = n;
    progLen[slot]++;
  }
  for (jint s = 0; s < numSegments; s++) {
    sq[s].filter = progFirst[(size_t)s * nProg];
    sq[s].num_filter_nodes = progLen[(size_t)s * nProg];
    sq[s].agg_filters = numAggFilters > 0 ? &progFirst[(size_t)s * nProg + 1] : NULL;
    sq[s].agg_filter_nodes = numAggFilters > 0 ? &progLen[(size_t)s * nProg + 1] : NULL;
  }
  const jsize nG = (*env)->GetArrayLength(env, groupBy), nA = (*env)->GetArrayLength(env, aggOps);
  const char** gb = (const char**)calloc((size_t)(nG > 0 ? nG : 1), sizeof(char*));
  for (jsize j = 0; j < nG; j++) gb[j] = (*env)->GetStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, groupBy, j), NULL);
  jint* ops = (*env)->GetIntArrayElements(env, aggOps, NULL);
  jint* fof = numAggFilters > 0 ? (*env)->GetIntArrayElements(env, aggFilterOf, NULL) : NULL;
  pb_aggregation_desc* aggs = (pb_aggregation_desc*)calloc((size_t)nA, sizeof(pb_aggregation_desc));
  for (jsize a = 0; a < nA; a++) {
    jstring c = (jstring)(*env)->GetObjectArrayElement(env, aggCols, a);
    aggs[a].op = ops[a];
    aggs[a].column = c ? (*env)->GetStringUTFChars(env, c, NULL) : NULL;     /* NULL for COUNT(*) */
  }
  pb_query_desc d;
  memset(&d, 0, sizeof d);
  d.num_group_by = (int32_t)nG; d.group_by_columns = gb;
  d.num_aggregations = (int32_t)nA; d.aggregations = aggs;
  d.num_groups_limit = numGroupsLimit; d.max_initial_result_holder_capacity = maxInitCapacity;
  d.flags = (uint32_t)flags;
  d.num_agg_filters = numAggFilters; d.agg_filter_of = (const int32_t*)fof;
  pb_result_handle r = NULL;
  int rc = pb_query_execute((pb_segment_group_handle)(intptr_t)group, sq, &d, &r);
  for (jsize a = 0; a < nA; a++) {
    jstring c = (jstring)(*env)->GetObjectArrayElement(env, aggCols, a);
    if (c && aggs[a].column) (*env)->ReleaseStringUTFChars(env, c, aggs[a].column);
  }
  for (jsize j = 0; j < nG; j++) (*env)->ReleaseStringUTFChars(env, (jstring)(*env)->GetObjectArrayElement(env, groupBy, j), gb[j]);
  if (fof) (*env)->ReleaseIntArrayElements(env, aggFilterOf, fof, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, aggOps, ops, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, rawPool, raws, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, idPool, ids, JNI_ABORT);
  (*env)->ReleaseDoubleArrayElements(env, nodeDoubles, nd, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, nodeLongs, nl, JNI_ABORT);
  (*env)->ReleaseIntArrayElements(env, nodeInts, ni, JNI_ABORT);
  free(aggs); free(gb); free(progLen); free(progFirst); free(sq); free(nodes);
  if (rc != PB_OK) { throw_last(env); return 0; }   /* PB_ERR_UNSUPPORTED here means the eligibility check and the engine disagree */
  return (jlong)(intptr_t)r;
}

JNIEXPORT void JNICALL Java_org_apache_pinot_b200_Native_releaseGroup(JNIEnv* env, jclass cls, jlong g) {
  pb_segment_group_release((pb_segment_group_handle)(intptr_t)g);
}

/* long[] resultStats(long result, int table): numDocsScanned, numEntriesScannedInFilter, numEntriesScannedPostFilter,
 * numTotalDocs, numGroupsLimitReached (ExecutionStatistics, CTR/operator/ExecutionStatistics.java:28-65) */
JNIEXPORT jlongArray JNICALL Java_org_apache_pinot_b200_Native_resultStats(JNIEnv* env, jclass cls, jlong r, jint table) {
  const pb_exec_stats* st = pb_result_stats((pb_result_handle)(intptr_t)r, table);
  jlongArray out = (*env)->NewLongArray(env, 5);
  if (!st) { throw_last(env); return out; }
  jlong v[5] = { st->num_docs_scanned, st->num_entries_scanned_in_filter, st->num_entries_scanned_post_filter, st->num_total_docs,
                 st->num_groups_limit_reached };
  (*env)->SetLongArrayRegion(env, out, 0, 5, v);
  return out;
}

/* ByteBuffer resultGroupKeyValues(long result, int table, int groupByColumn): decoded key values, native-endian, fixed width
 * (INT 4, LONG 8, FLOAT 4, DOUBLE 8, STRING lengthOfEachEntry zero-padded) — GroupKeyGenerator.GroupKey._keys */
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultGroupKeyValues(JNIEnv* env, jclass cls, jlong r, jint table, jint gb) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  int32_t type = 0, eb = 0;
  const void* p = pb_result_group_key_values(h, table, gb, &type, &eb);
  if (!p) { throw_last(env); return NULL; }
  return (*env)->NewDirectByteBuffer(env, (void*)p, (jlong)eb * pb_result_num_groups(h, table));
}

/* DISTINCTCOUNT value sets (the intermediate result the combine layer merges): offsets[numGroups + 1] into dictIds[] */
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultDistinctOffsets(JNIEnv* env, jclass cls, jlong r, jint table, jint agg) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  const int64_t* p = pb_result_distinct_offsets(h, table, agg);
  if (!p) { throw_last(env); return NULL; }
  return (*env)->NewDirectByteBuffer(env, (void*)p, 8 * (pb_result_num_groups(h, table) + 1));
}
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultDistinctDictIds(JNIEnv* env, jclass cls, jlong r, jint table, jint agg) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  const int64_t* off = pb_result_distinct_offsets(h, table, agg);
  const int32_t* p = pb_result_distinct_dict_ids(h, table, agg);
  if (!off || !p) { throw_last(env); return NULL; }
  return (*env)->NewDirectByteBuffer(env, (void*)p, 4 * off[pb_result_num_groups(h, table)]);
}

JNIEXPORT jlong JNICALL Java_org_apache_pinot_b200_Native_resultNumGroups(JNIEnv* env, jclass cls, jlong r, jint table) {
  return (jlong)pb_result_num_groups((pb_result_handle)(intptr_t)r, table);
}
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultDoubles(JNIEnv* env, jclass cls, jlong r, jint table, jint agg) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  return (*env)->NewDirectByteBuffer(env, (void*)pb_result_double(h, table, agg), 8 * pb_result_num_groups(h, table));
}
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultLongs(JNIEnv* env, jclass cls, jlong r, jint table, jint agg) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  return (*env)->NewDirectByteBuffer(env, (void*)pb_result_long(h, table, agg), 8 * pb_result_num_groups(h, table));
}
JNIEXPORT jobject JNICALL Java_org_apache_pinot_b200_Native_resultGroupDictIds(JNIEnv* env, jclass cls, jlong r, jint table, jint gb) {
  pb_result_handle h = (pb_result_handle)(intptr_t)r;
  return (*env)->NewDirectByteBuffer(env, (void*)pb_result_group_dict_ids(h, table, gb), 4 * pb_result_num_groups(h, table));
}
JNIEXPORT void JNICALL Java_org_apache_pinot_b200_Native_freeResult(JNIEnv* env, jclass cls, jlong r) {
  pb_result_free((pb_result_handle)(intptr_t)r);
}
#endif /* PB_WITH_JNI */
