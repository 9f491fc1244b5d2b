// The subset of caffe.proto this path reads (ref: src/caffe/proto/caffe.proto:75-173 solver, :226-330
// LayerParameter, :428-433 concat, :562-620 video_sampled_shots_data_param, :697-745 dropout/eltwise/sum,
// :831-868 inner_product / max_margin_loss), as accessor classes over a parsed protobuf *text* tree, with
// the generated-code accessor names (num_output(), coeff_size(), coeff(i), has_xxx()) and the .proto
// defaults, so layer bodies read their parameters exactly as in the reference.
#pragma once
#include <map>
#include "caffe/common.hpp"

namespace caffe {

class PbMsg;
// `floats`: a packed repeated float field (BlobProto.data / .diff) kept as one array instead of one scalar per element
struct PbField { string key; string scalar; shared_ptr<PbMsg> msg; shared_ptr<vector<float> > floats; };
class PbMsg {
 public:
  vector<PbField> fields;
  int count(const string& key) const;
  const PbField* nth(const string& key, int i) const;
  bool has(const string& key) const { return count(key) > 0; }
  string str(const string& key, const string& dflt = "", int i = 0) const;
  double num(const string& key, double dflt, int i = 0) const;
  bool boolean(const string& key, bool dflt) const;
  shared_ptr<PbMsg> sub(const string& key, int i = 0) const;   // empty message if absent
  void set_scalar(const string& key, const string& v);
  void add_scalar(const string& key, const string& v);
};
// protobuf text format: `key: value`, `key { ... }`, `key: { ... }`, '#' comments, quoted strings
shared_ptr<PbMsg> ParseTextFormat(const string& text);
string PrintTextFormat(const PbMsg& m);
string ReadFileOrDie(const string& path);

// protobuf *binary* wire format for the reference's message types (ref: util/io.cpp:49-67 Read/WriteProtoFrom/ToBinaryFile;
// schema = caffe/proto/caffe_schema.inc).  `type` names the root message ("NetParameter", "SolverState", "BlobProto" ...).
// Fields are written in field-number order (repeated ones in stored order), enums by number, unknown fields skipped on read.
string SerializeBinary(const PbMsg& m, const string& type);
shared_ptr<PbMsg> ParseBinary(const string& bytes, const string& type);
void WriteProtoToBinaryFile(const PbMsg& m, const string& type, const string& path);
shared_ptr<PbMsg> ReadProtoFromBinaryFile(const string& path, const string& type);

enum LayerParameter_LayerType {   // values as in caffe.proto:236-302
  LayerParameter_LayerType_NONE = 0, LayerParameter_LayerType_CONCAT = 3, LayerParameter_LayerType_DROPOUT = 6,
  LayerParameter_LayerType_FLATTEN = 8, LayerParameter_LayerType_INNER_PRODUCT = 14, LayerParameter_LayerType_RELU = 18,
  LayerParameter_LayerType_SPLIT = 22, LayerParameter_LayerType_ELTWISE = 25, LayerParameter_LayerType_SLICE = 33,
  LayerParameter_LayerType_NORMALIZATION = 41, LayerParameter_LayerType_ID_TO_WEIGHT_MAPPING = 42, LayerParameter_LayerType_MAX_MARGIN_LOSS = 43,
  LayerParameter_LayerType_SUM = 44, LayerParameter_LayerType_RETRIEVAL_STATS = 45,
  LayerParameter_LayerType_VIDEO_SHOT_WINDOW_TEST_DATA = 48, LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA = 49
};
LayerParameter_LayerType LayerTypeFromName(const string& name);
const char* LayerTypeName(LayerParameter_LayerType t);

enum EltwiseParameter_EltwiseOp { EltwiseParameter_EltwiseOp_PROD = 0, EltwiseParameter_EltwiseOp_SUM = 1, EltwiseParameter_EltwiseOp_MAX = 2 };
enum MaxMarginLossParameter_Norm { MaxMarginLossParameter_Norm_L1 = 1, MaxMarginLossParameter_Norm_L2 = 2 };
enum VideoSampledShotsDataParameter_ContextType {   // values as in caffe.proto:598-604 (= VV_CONTEXT_*)
  VideoSampledShotsDataParameter_CONTEXT_PAIRWISE = 0, VideoSampledShotsDataParameter_CONTEXT_WINDOW = 1,
  VideoSampledShotsDataParameter_CONTEXT_PAST = 2, VideoSampledShotsDataParameter_CONTEXT_PAST_CONTINUOUS = 3,
  VideoSampledShotsDataParameter_CONTEXT_PAST_CONTINUOUS_FIXED = 4 };

struct ParamBase { shared_ptr<PbMsg> m; explicit ParamBase(shared_ptr<PbMsg> p) : m(p ? p : std::make_shared<PbMsg>()) {} };

struct FillerParameter : ParamBase {
  using ParamBase::ParamBase;
  string type() const { return m->str("type", "constant"); }
  float value() const { return float(m->num("value", 0)); }
  float min() const { return float(m->num("min", 0)); }
  float max() const { return float(m->num("max", 1)); }
  float mean() const { return float(m->num("mean", 0)); }
  float std() const { return float(m->num("std", 1)); }
};
struct InnerProductParameter : ParamBase {
  using ParamBase::ParamBase;
  int num_output() const { return int(m->num("num_output", 0)); }
  bool bias_term() const { return m->boolean("bias_term", true); }
  FillerParameter weight_filler() const { return FillerParameter(m->sub("weight_filler")); }
  FillerParameter bias_filler() const { return FillerParameter(m->sub("bias_filler")); }
  float regularization() const { return float(m->num("regularization", 0)); }   // fork-added, proto:836
};
struct IdToWeightMappingParameter : ParamBase {     // caffe.proto:824-828
  using ParamBase::ParamBase;
  int num_output() const { return int(m->num("num_output", 0)); }
  int max_ids() const { return int(m->num("max_ids", 0)); }
  FillerParameter weight_filler() const { return FillerParameter(m->sub("weight_filler")); }
};
struct EltwiseParameter : ParamBase {
  using ParamBase::ParamBase;
  EltwiseParameter_EltwiseOp operation() const;
  int coeff_size() const { return m->count("coeff"); }
  float coeff(int i) const { return float(m->num("coeff", 0, i)); }
  bool stable_prod_grad() const { return m->boolean("stable_prod_grad", true); }
};
struct SumParameter : ParamBase { using ParamBase::ParamBase; float num_output() const { return float(m->num("num_output", 1)); } };
struct MaxMarginLossParameter : ParamBase {
  using ParamBase::ParamBase;
  MaxMarginLossParameter_Norm norm() const { return m->str("norm", "L1") == "L2" ? MaxMarginLossParameter_Norm_L2 : MaxMarginLossParameter_Norm_L1; }
  string id_to_weight_file() const { return m->str("id_to_weight_file", ""); }
  bool use_direct_weight() const { return m->boolean("use_direct_weight", false); }
  float margin() const { return float(m->num("margin", 1.0)); }
};
struct DropoutParameter : ParamBase { using ParamBase::ParamBase; float dropout_ratio() const { return float(m->num("dropout_ratio", 0.5)); } };
struct ReLUParameter : ParamBase { using ParamBase::ParamBase; float negative_slope() const { return float(m->num("negative_slope", 0)); } };
struct SliceParameter : ParamBase {
  using ParamBase::ParamBase;
  unsigned slice_dim() const { return unsigned(m->num("slice_dim", 1)); }
  int slice_point_size() const { return m->count("slice_point"); }
  unsigned slice_point(int i) const { return unsigned(m->num("slice_point", 0, i)); }
};
struct ConcatParameter : ParamBase { using ParamBase::ParamBase; unsigned concat_dim() const { return unsigned(m->num("concat_dim", 1)); } };
struct VideoSampledShotsDataParameter : ParamBase {
  using ParamBase::ParamBase;
  string source() const { return m->str("source", ""); }
  string negative_dataset() const { return m->str("negative_dataset", ""); }
  int batch_size() const { return int(m->num("batch_size", 0)); }
  int rand_skip() const { return int(m->num("rand_skip", 0)); }
  int num_negative_samples() const { return int(m->num("num_negative_samples", 0)); }
  int max_buffer_size() const { return int(m->num("max_buffer_size", 5000)); }
  int negative_swap_percentage() const { return int(m->num("negative_swap_percentage", 50)); }
  int max_same_video_negs() const { return int(m->num("max_same_video_negs", 0)); }
  int context_size() const { return int(m->num("context_size", 5)); }
  VideoSampledShotsDataParameter_ContextType context_type() const;
};
// ref: caffe.proto:538-559 (TEST-phase data layer) and :956-967
struct VideoShotWindowTestDataParameter : ParamBase {
  using ParamBase::ParamBase;
  string source() const { return m->str("source", ""); }
  int batch_size() const { return int(m->num("batch_size", 0)); }
  bool include_positives() const { return m->boolean("include_positives", true); }
  bool include_negatives() const { return m->boolean("include_negatives", true); }
};
struct RetrievalStatsParameter : ParamBase {
  using ParamBase::ParamBase;
  string id_to_class_file() const { return m->str("id_to_class_file", ""); }
  string stats_output_file() const { return m->str("stats_output_file", ""); }
  bool exclude_same_video_shots() const { return m->boolean("exclude_same_video_shots", true); }
  bool video_level_retrieval() const { return m->boolean("video_level_retrieval", false); }
  int max_num_videos() const { return int(m->num("max_num_videos", 0)); }
};
struct NetStateRule : ParamBase { using ParamBase::ParamBase; bool has_phase() const { return m->has("phase"); } Caffe::Phase phase() const { return m->str("phase") == "TEST" ? Caffe::TEST : Caffe::TRAIN; } };

struct LayerParameter : ParamBase {
  using ParamBase::ParamBase;
  LayerParameter() : ParamBase(nullptr) {}
  string name() const { return m->str("name"); }
  LayerParameter_LayerType type() const { return LayerTypeFromName(m->str("type")); }
  int bottom_size() const { return m->count("bottom"); }
  string bottom(int i) const { return m->str("bottom", "", i); }
  int top_size() const { return m->count("top"); }
  string top(int i) const { return m->str("top", "", i); }
  int blobs_lr_size() const { return m->count("blobs_lr"); }
  float blobs_lr(int i) const { return float(m->num("blobs_lr", 1, i)); }
  int weight_decay_size() const { return m->count("weight_decay"); }
  float weight_decay(int i) const { return float(m->num("weight_decay", 1, i)); }
  int loss_weight_size() const { return m->count("loss_weight"); }
  float loss_weight(int i) const { return float(m->num("loss_weight", 0, i)); }
  int include_size() const { return m->count("include"); }
  NetStateRule include(int i) const { return NetStateRule(m->sub("include", i)); }
  int blobs_size() const { return 0; }   // serialized blobs arrive through CopyTrainedLayersFrom, not text
  InnerProductParameter inner_product_param() const { return InnerProductParameter(m->sub("inner_product_param")); }
  EltwiseParameter eltwise_param() const { return EltwiseParameter(m->sub("eltwise_param")); }
  IdToWeightMappingParameter id_to_weight_mapping_param() const { return IdToWeightMappingParameter(m->sub("id_to_weight_mapping_param")); }
  SumParameter sum_param() const { return SumParameter(m->sub("sum_param")); }
  MaxMarginLossParameter max_margin_loss_param() const { return MaxMarginLossParameter(m->sub("max_margin_loss_param")); }
  DropoutParameter dropout_param() const { return DropoutParameter(m->sub("dropout_param")); }
  ReLUParameter relu_param() const { return ReLUParameter(m->sub("relu_param")); }
  SliceParameter slice_param() const { return SliceParameter(m->sub("slice_param")); }
  ConcatParameter concat_param() const { return ConcatParameter(m->sub("concat_param")); }
  VideoSampledShotsDataParameter video_sampled_shots_data_param() const { return VideoSampledShotsDataParameter(m->sub("video_sampled_shots_data_param")); }
  VideoShotWindowTestDataParameter video_shot_window_test_data_param() const { return VideoShotWindowTestDataParameter(m->sub("video_shot_window_test_data_param")); }
  RetrievalStatsParameter retrieval_stats_param() const { return RetrievalStatsParameter(m->sub("retrieval_stats_param")); }
  // builders used by InsertSplits and tests
  void set_name(const string& v) { m->set_scalar("name", v); }
  void set_type(LayerParameter_LayerType t) { m->set_scalar("type", LayerTypeName(t)); }
  void add_bottom(const string& v) { m->add_scalar("bottom", v); }
  void add_top(const string& v) { m->add_scalar("top", v); }
  void add_loss_weight(float v) { m->add_scalar("loss_weight", std::to_string(v)); }
};

struct NetParameter : ParamBase {
  using ParamBase::ParamBase;
  string name() const { return m->str("name"); }
  int layers_size() const { return m->count("layers"); }
  LayerParameter layers(int i) const { return LayerParameter(m->sub("layers", i)); }
};
struct SolverParameter : ParamBase {
  using ParamBase::ParamBase;
  string net() const { return m->str("net"); }
  string train_net() const { return m->str("train_net"); }
  float base_lr() const { return float(m->num("base_lr", 0.01)); }
  string lr_policy() const { return m->str("lr_policy", "fixed"); }
  float gamma() const { return float(m->num("gamma", 0)); }
  float power() const { return float(m->num("power", 0)); }
  int stepsize() const { return int(m->num("stepsize", 1)); }
  float momentum() const { return float(m->num("momentum", 0)); }
  float weight_decay() const { return float(m->num("weight_decay", 0)); }
  string regularization_type() const { return m->str("regularization_type", "L2"); }
  int max_iter() const { return int(m->num("max_iter", 0)); }
  int display() const { return int(m->num("display", 0)); }
  int snapshot() const { return int(m->num("snapshot", 0)); }
  string snapshot_prefix() const { return m->str("snapshot_prefix", ""); }
  bool snapshot_diff() const { return m->boolean("snapshot_diff", false); }
  bool snapshot_after_train() const { return m->boolean("snapshot_after_train", true); }
  long random_seed() const { return long(m->num("random_seed", -1)); }
  int test_interval() const { return int(m->num("test_interval", 0)); }
  int test_iter_size() const { return m->count("test_iter"); }
  int test_iter(int i) const { return int(m->num("test_iter", 0, i)); }
  int test_net_size() const { return m->count("test_net"); }
  string test_net(int i) const { return m->str("test_net", "", i); }
  bool test_initialization() const { return m->boolean("test_initialization", true); }
  bool test_compute_loss() const { return m->boolean("test_compute_loss", false); }
  string solver_mode() const { return m->str("solver_mode", "GPU"); }
  int device_id() const { return int(m->num("device_id", 0)); }
  bool debug_info() const { return m->boolean("debug_info", false); }
};

NetParameter ReadNetParamsFromTextFileOrDie(const string& path);
SolverParameter ReadSolverParamsFromTextFileOrDie(const string& path);

}  // namespace caffe
