// pb_internal.h — read-only views of staged segments shared between the device runtime (pb_engine.cu) and the
// host planning layer (host/pb_host.cpp).  Not part of the public ABI.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/pinot_b200.h"

struct PbColumnView {
  std::string name;
  int type = 0, has_dict = 0, is_sorted = 0, card = 0, bits = 0, entry_bytes = 0;
  const uint8_t* dict = nullptr;          // big-endian dictionary bytes (host copy)
  const int32_t* sorted_pairs = nullptr;  // sorted column: little-endian (start,end) inclusive docId pairs
  bool has_inverted = false;
  const uint8_t* null_vector = nullptr;   // DataSource.getNullValueVector(): RoaringBitmap of the null docIds (caller's buffer)
  uint64_t null_vector_len = 0;
};
struct PbSegmentView {
  std::string name;
  int num_docs = 0;
  std::vector<PbColumnView> cols;
};
int pbi_segment_view(pb_segment_handle seg, PbSegmentView* out);
int pbi_group_segments(pb_segment_group_handle g, std::vector<pb_segment_handle>* out);
int pbi_fail(int code, const char* msg);
// plan cache, host side: the key of the unlowered query lets a repeated query skip the per-segment lowering as well
void pbi_set_pending_host_key(const std::string& key);
int pbi_plan_replay(pb_segment_group_handle g, const std::string& host_key, const pb_query_desc* q, pb_result_handle* out);
