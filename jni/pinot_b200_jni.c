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
 * ByteBuffer[] inv) */
JNIEXPORT jlong JNICALL Java_org_apache_pinot_b200_Native_stageSegment(JNIEnv* env, jclass cls, jstring name, jint numDocs,
    jobjectArray colNames, jintArray meta, jobjectArray fwd, jobjectArray dict, jobjectArray inv) {
  jsize n = (*env)->GetArrayLength(env, colNames);
  pb_column_desc* cols = (pb_column_desc*)calloc((size_t)n, sizeof(pb_column_desc));
  jint* m = (*env)->GetIntArrayElements(env, meta, NULL);
  const char** names = (const char**)calloc((size_t)n, sizeof(char*));
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, colNames, i);
    names[i] = (*env)->GetStringUTFChars(env, s, NULL);
    cols[i].name = names[i];
    cols[i].stored_type = m[6 * i]; cols[i].has_dictionary = m[6 * i + 1]; cols[i].is_sorted = m[6 * i + 2];
    cols[i].cardinality = m[6 * i + 3]; cols[i].bits_per_element = m[6 * i + 4]; cols[i].dict_entry_bytes = m[6 * i + 5];
    jobject b = (*env)->GetObjectArrayElement(env, fwd, i);
    cols[i].forward_index = (*env)->GetDirectBufferAddress(env, b);
    cols[i].forward_index_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b);
    b = (*env)->GetObjectArrayElement(env, dict, i);
    if (b) { cols[i].dictionary = (*env)->GetDirectBufferAddress(env, b); cols[i].dictionary_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b); }
    b = (*env)->GetObjectArrayElement(env, inv, i);
    if (b) { cols[i].inverted_index = (*env)->GetDirectBufferAddress(env, b); cols[i].inverted_index_len = (uint64_t)(*env)->GetDirectBufferCapacity(env, b); }
  }
  const char* sname = (*env)->GetStringUTFChars(env, name, NULL);
  pb_segment_desc d = { sname, numDocs, (int32_t)n, cols };
  pb_segment_handle h = NULL;
  int rc = pb_segment_stage(&d, 0, &h);
  (*env)->ReleaseStringUTFChars(env, name, sname);
  (*env)->ReleaseIntArrayElements(env, meta, m, JNI_ABORT);
  free(names); free(cols);
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

/* long execute(long group, int[] nodeInts /\* 8 ints per node: segment, kind, column, numChildren, exclusive, numIds,
 * dloIncl, dhiIncl *\/, long[] nodeLongs /\* lo, hi per node *\/, double[] nodeDoubles /\* dlo, dhi per node *\/,
 * int[][] nodeIds, String[] groupBy, int[] aggOps, String[] aggCols, int numGroupsLimit, int maxInitCapacity, int flags) —
 * flattening / unflattening is mechanical and elided here for brevity of the shim; see Native.java for the layout. */

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
