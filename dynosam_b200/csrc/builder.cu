// builder.cu -- SURVEY.md 8f-3: graph construction straight into the flat SoA blocks of the C ABI.
//
// The reference walks its Map (frames -> observed tracklets / objects) and allocates one heap factor per observation
// (Formulation<MAP>::updateStaticObservations / updateDynamicObservations, dynosam/include/dynosam/backend/
// Formulation-impl.hpp:552-897; the hybrid rules in src/backend/rgbd/HybridEstimator.cc:573-830), only for the solver to
// walk them again.  This builder takes the same information -- per frame: camera pose estimate, odometry, static and
// dynamic point observations, front-end object motions -- and applies the formulation's topology rules once, emitting the
// variable arrays and one homogeneous block per factor type, ready for dynoba_set_variables / dynoba_add_factors:
//   * camera chain: PriorFactor(X_first, sigma 1e-6) + BetweenFactor odometry (VisionImuBackendModule.hpp:168-243)
//   * static points: added once seen min_static_obs times, initial value X_k z of the first observation, one
//     PoseToPointFactor per observation (Formulation-impl.hpp:145-212)
//   * HYBRID: object key-frame e = first frame of a visibility segment (a new key-frame once the object was unseen for more
//     than keyframe_gap frames, HybridEstimator.cc:985-998); motion variable e_H_k per (object, frame); one m_L per tracklet
//     initialised by HybridObjectMotion::projectToObject3 at its first observation; HybridMotionFactor(X_k, e_H_k, m_L; z, L_e);
//     PriorFactor(H_e = I, 1e-6) at the key-frame (:744-746); three-motion HybridSmoothingFactor inside a segment (:800-802);
//     L_e = centroid of the key-frame's points with identity rotation unless given (:867-875).
//   * WCME (formulation 1; WorldMotionEstimator.cc:151-351): one point per (tracklet, frame) initialised X_k z, a PoseToPointFactor
//     per observation, LandmarkMotionTernaryFactor(m_prev, m_k, H_k) between consecutive observations of a tracklet, a motion
//     variable H_k per (object, frame that closes a pair) initialised with the front end's translation and IDENTITY rotation
//     (:297-303), BetweenFactor(H_k-1, H_k, I) smoothing between consecutive frames (:309-349).
//   * WCPE (formulation 2; WorldPoseEstimator.cc:89-315): the same points, LandmarkMotionPoseFactor(m_prev, m_k, L_prev, L_k), an
//     object pose L_k per (object, frame) initialised motion * L_k-1 when the front end gave a motion and L_k-1 exists, else the
//     centroid of the object's points at k with identity rotation (:205-232), LandmarkPoseSmoothingFactor over three consecutive
//     frames (:259-306); no prior on the object poses.
// Host code only (no kernel): it lives in libdynoba.so because its output is the library's input.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dynoba.h"
#include "se3.cuh"

using namespace dynoba;

namespace {

struct Obs { int32_t frame; int64_t tracklet; int32_t object; double z[3]; };
struct Block { int type = 0; std::vector<int32_t> idx, aux; std::vector<double> meas, sigma; int sigma_dim = 1; bool bcast = true; double k = 0; bool has_aux = false;
               int64_t n() const { static const int ar[11] = {1, 2, 2, 2, 3, 3, 3, 4, 3, 3, 2}; return (int64_t)idx.size()/ar[type]; } };

Pose pose_from(const double* p) { Pose P; for (int i = 0; i < 9; i++) P.R[i] = p[i]; for (int i = 0; i < 3; i++) P.t[i] = p[9 + i]; return P; }
void pose_to(const Pose& P, double* p) { for (int i = 0; i < 9; i++) p[i] = P.R[i]; for (int i = 0; i < 3; i++) p[9 + i] = P.t[i]; }
Pose identity_pose() { Pose P; for (int i = 0; i < 9; i++) P.R[i] = (i % 4 == 0) ? 1.0 : 0.0; P.t[0] = P.t[1] = P.t[2] = 0.0; return P; }
uint64_t sym(char c, uint64_t j) { return ((uint64_t)(unsigned char)c << 56) | j; }
uint64_t labeled(char c, uint64_t label, uint64_t j) { return ((uint64_t)(unsigned char)c << 56) | (label << 48) | j; }
uint64_t cantor(uint64_t k1, uint64_t k2) { return (k1 + k2)*(k1 + k2 + 1)/2 + k2; }

}  // namespace

struct dynoba_builder {
  dynoba_builder_params prm; std::string err; bool finalized = false;
  std::map<int32_t, std::pair<Pose, bool>> frames;             // frame -> (camera pose estimate, has odometry)
  std::map<int32_t, Pose> odom;                                  // frame -> relative pose from the previous frame
  std::vector<Obs> stat, dyn;
  std::map<std::pair<int32_t, int32_t>, Pose> motion_init, keyframe_pose;    // (object, frame)
  // products
  std::vector<double> pose, point, aux; std::vector<int32_t> order; std::vector<uint64_t> pose_keys, point_keys;
  std::vector<Block> blocks;
};

#define BARG(cond, msg) do { if (!(cond)) { if (b) b->err = msg; return DYNOBA_ERR_BAD_ARG; } } while (0)

extern "C" {

void dynoba_builder_default_params(dynoba_builder_params* p) {
  p->min_static_obs = 2; p->min_dynamic_obs = 3; p->keyframe_gap = 2;       // params: min_static_observations, min_dynamic_observations
  p->sigma_static = 0.2; p->sigma_dynamic = 0.2; p->huber_k = 1e-4;           // BackendDefinitions.cc:124-196 flag defaults
  const double od[6] = {0.02, 0.02, 0.02, 0.01, 0.01, 0.01}, sm[6] = {0.01, 0.01, 0.01, 0.1, 0.1, 0.1};
  for (int i = 0; i < 6; i++) { p->odometry_sigma[i] = od[i]; p->smoothing_sigma[i] = sm[i]; }
  p->prior_sigma = 1e-6;
  p->formulation = DYNOBA_FORMULATION_HYBRID; p->sigma_motion = 0.01;         // motion_ternary_factor_noise_sigma
  p->backtrack = 1;
}
int dynoba_builder_create(const dynoba_builder_params* p, dynoba_builder_handle* out) {
  if (!out) return DYNOBA_ERR_BAD_ARG;
  dynoba_builder* b = new dynoba_builder();
  if (p) b->prm = *p; else dynoba_builder_default_params(&b->prm);
  *out = b; return DYNOBA_OK;
}
int dynoba_builder_destroy(dynoba_builder_handle b) { if (!b) return DYNOBA_ERR_BAD_ARG; delete b; return DYNOBA_OK; }
const char* dynoba_builder_last_error(dynoba_builder_handle b) { return b ? b->err.c_str() : "null builder"; }

int dynoba_builder_add_frame(dynoba_builder_handle b, int32_t frame, const double* X, const double* odom_from_prev) {
  BARG(b && X && frame >= 0, "bad frame");
  b->frames[frame] = { pose_from(X), odom_from_prev != nullptr };
  if (odom_from_prev) b->odom[frame] = pose_from(odom_from_prev);
  b->finalized = false; return DYNOBA_OK;
}
int dynoba_builder_add_static(dynoba_builder_handle b, int32_t frame, int64_t n, const int64_t* tracklet, const double* z) {
  BARG(b && n >= 0 && (n == 0 || (tracklet && z)), "bad static observations");
  for (int64_t i = 0; i < n; i++) b->stat.push_back(Obs{ frame, tracklet[i], 0, { z[3*i], z[3*i + 1], z[3*i + 2] } });
  b->finalized = false; return DYNOBA_OK;
}
int dynoba_builder_add_dynamic(dynoba_builder_handle b, int32_t frame, int64_t n, const int64_t* tracklet, const int32_t* object, const double* z) {
  BARG(b && n >= 0 && (n == 0 || (tracklet && object && z)), "bad dynamic observations");
  for (int64_t i = 0; i < n; i++) { BARG(object[i] > 0, "object ids start at 1 (0 is the background)"); b->dyn.push_back(Obs{ frame, tracklet[i], object[i], { z[3*i], z[3*i + 1], z[3*i + 2] } }); }
  b->finalized = false; return DYNOBA_OK;
}
int dynoba_builder_set_motion_init(dynoba_builder_handle b, int32_t object, int32_t frame, const double* H) {
  BARG(b && H, "null"); b->motion_init[{object, frame}] = pose_from(H); b->finalized = false; return DYNOBA_OK;
}
int dynoba_builder_set_keyframe_pose(dynoba_builder_handle b, int32_t object, int32_t keyframe, const double* L_e) {
  BARG(b && L_e, "null"); b->keyframe_pose[{object, keyframe}] = pose_from(L_e); b->finalized = false; return DYNOBA_OK;
}

int dynoba_builder_finalize(dynoba_builder_handle b) {
  BARG(b, "null builder");
  if (b->finalized) return DYNOBA_OK;
  const dynoba_builder_params& P = b->prm;
  b->pose.clear(); b->point.clear(); b->aux.clear(); b->order.clear(); b->pose_keys.clear(); b->point_keys.clear(); b->blocks.clear();
  BARG(!b->frames.empty(), "no frames");
  // ---- camera poses, in frame order
  std::map<int32_t, int32_t> cam_index;
  for (auto& kv : b->frames) {
    cam_index[kv.first] = (int32_t)cam_index.size();
    double p[12]; pose_to(kv.second.first, p); b->pose.insert(b->pose.end(), p, p + 12);
    b->order.push_back(kv.first); b->pose_keys.push_back(sym('X', (uint64_t)kv.first));
  }
  auto cam = [&](int32_t f) -> int32_t { auto it = cam_index.find(f); return it == cam_index.end() ? -1 : it->second; };
  // ---- static landmarks: tracklets in order of first appearance (frame-major, insertion order inside a frame)
  auto by_frame = [](const Obs& a, const Obs& c) { return a.frame < c.frame; };
  std::stable_sort(b->stat.begin(), b->stat.end(), by_frame);
  std::stable_sort(b->dyn.begin(), b->dyn.end(), by_frame);
  // ---- which observations enter the graph.  A tracklet needs min_*_obs observations.  With backtrack (the batch graph of
  // SURVEY 8d; UpdateObservationParams::do_backtrack = true, ParallelHybridBackendModule.cc:427) all of them are added; without
  // (RegularBackendModule.cc:139,197) the tracklet enters at the frame its count reaches the minimum and only what that update
  // adds is kept: static -- that frame's observation on (Formulation-impl.hpp:194-199); dynamic -- the pair (previous, that
  // frame) on (:703-720), i.e. the first min_dynamic_obs - 2 observations never enter.
  std::vector<Obs> stat, dyn;
  {
    std::map<std::pair<int32_t, int64_t>, int> count, seen;
    for (auto& o : b->stat) count[{0, o.tracklet}]++;
    for (auto& o : b->dyn) count[{o.object, o.tracklet}]++;
    const int min_dyn = P.formulation == DYNOBA_FORMULATION_HYBRID ? P.min_dynamic_obs : std::max(P.min_dynamic_obs, 2);   // a motion factor needs a pair
    const int start_s = P.backtrack ? 0 : std::max(P.min_static_obs - 1, 0), start_d = P.backtrack ? 0 : std::max(min_dyn - 2, 0);
    for (auto& o : b->stat) if (count[{0, o.tracklet}] >= P.min_static_obs && seen[{0, o.tracklet}]++ >= start_s) stat.push_back(o);
    for (auto& o : b->dyn) if (count[{o.object, o.tracklet}] >= min_dyn && seen[{o.object, o.tracklet}]++ >= start_d) dyn.push_back(o);
  }
  // ---- static landmarks: tracklets in order of first appearance (frame-major, insertion order inside a frame)
  {
    std::map<int64_t, std::vector<size_t>> tracks; std::vector<int64_t> first_seen;
    for (size_t i = 0; i < stat.size(); i++) { BARG(cam(stat[i].frame) >= 0, "static observation in an unknown frame");
      auto& t = tracks[stat[i].tracklet]; if (t.empty()) first_seen.push_back(stat[i].tracklet); t.push_back(i); }
    Block blk; blk.type = DYNOBA_POSE2POINT3; blk.sigma = { P.sigma_static }; blk.k = P.huber_k;
    for (int64_t tid : first_seen) {
      auto& t = tracks[tid];
      const int32_t pi = (int32_t)(b->point.size()/3);
      const Obs& o0 = stat[t[0]];
      double w[3]; se3_transform_from(b->frames[o0.frame].first, o0.z, w);          // initial value: X_k z of the first observation that enters
      b->point.insert(b->point.end(), w, w + 3); b->point_keys.push_back(sym('l', (uint64_t)tid));
      for (size_t i : t) { const Obs& o = stat[i]; blk.idx.push_back(cam(o.frame)); blk.idx.push_back(pi); blk.meas.insert(blk.meas.end(), o.z, o.z + 3); }
    }
    if (blk.n()) b->blocks.push_back(std::move(blk));
  }
  if (P.formulation != DYNOBA_FORMULATION_HYBRID) {
    // ---- world-centric formulations: a point per (tracklet, observation), chained by motion factors
    BARG(P.formulation == DYNOBA_FORMULATION_WCME || P.formulation == DYNOBA_FORMULATION_WCPE, "unknown formulation");
    const bool wcpe = P.formulation == DYNOBA_FORMULATION_WCPE;
    const int32_t n_cam = (int32_t)cam_index.size();
    std::map<std::pair<int32_t, int64_t>, std::vector<size_t>> tracks; std::vector<std::pair<int32_t, int64_t>> first_seen;
    for (size_t i = 0; i < dyn.size(); i++) { BARG(cam(dyn[i].frame) >= 0, "dynamic observation in an unknown frame");
      auto key = std::make_pair(dyn[i].object, dyn[i].tracklet);
      auto& t = tracks[key]; if (t.empty()) first_seen.push_back(key); t.push_back(i); }
    std::stable_sort(first_seen.begin(), first_seen.end(), [](const std::pair<int32_t, int64_t>& a, const std::pair<int32_t, int64_t>& c) { return a.first < c.first; });
    // (object, frame) -> pose-like variable: WCME the frame that closes a pair, WCPE every observed frame of a kept tracklet
    std::map<std::pair<int32_t, int32_t>, int32_t> var_index;
    std::map<std::pair<int32_t, int32_t>, std::pair<std::array<double, 3>, int>> centroid;     // world points of the kept tracklets
    for (auto& key : first_seen) {
      auto& t = tracks[key];
      for (size_t k = 0; k < t.size(); k++) { const Obs& o = dyn[t[k]];
        if (k > 0) BARG(o.frame > dyn[t[k-1]].frame, "a tracklet is observed twice in one frame");
        if (wcpe || k > 0) var_index[{o.object, o.frame}] = -1;
        double w[3]; se3_transform_from(b->frames[o.frame].first, o.z, w);
        auto& c = centroid[{o.object, o.frame}]; for (int a = 0; a < 3; a++) c.first[a] += w[a]; c.second++; }
    }
    for (auto& kv : var_index) {                                       // std::map order: object-major, frames ascending
      const int32_t obj = kv.first.first, fr = kv.first.second;
      kv.second = (int32_t)(b->pose.size()/12);
      Pose V = identity_pose();
      auto mi = b->motion_init.find({obj, fr});
      if (!wcpe) { if (mi != b->motion_init.end()) for (int a = 0; a < 3; a++) V.t[a] = mi->second.t[a]; }
      else {
        auto given = b->keyframe_pose.find({obj, fr}); auto prev = var_index.find({obj, fr - 1});
        if (given != b->keyframe_pose.end()) V = given->second;
        else if (mi != b->motion_init.end() && prev != var_index.end() && prev->second >= 0) se3_compose(mi->second, pose_from(&b->pose[(size_t)12*prev->second]), V);
        else { auto& c = centroid[{obj, fr}]; for (int a = 0; a < 3; a++) V.t[a] = c.first[a]/std::max(c.second, 1); }
      }
      double p[12]; pose_to(V, p); b->pose.insert(b->pose.end(), p, p + 12);
      b->order.push_back(fr); b->pose_keys.push_back(labeled(wcpe ? 'L' : 'H', (uint64_t)('0' + obj), (uint64_t)fr));
    }
    Block ptp; ptp.type = DYNOBA_POSE2POINT3; ptp.sigma = { P.sigma_dynamic }; ptp.k = P.huber_k;
    Block mot; mot.type = wcpe ? DYNOBA_MOTIONPOSE3 : DYNOBA_TERNARY3; mot.sigma = { P.sigma_motion }; mot.k = P.huber_k;
    for (auto& key : first_seen) {
      auto& t = tracks[key];
      for (size_t k = 0; k < t.size(); k++) { const Obs& o = dyn[t[k]];
        const int32_t pi = (int32_t)(b->point.size()/3);
        double w[3]; se3_transform_from(b->frames[o.frame].first, o.z, w);                 // X_k z (dynamicPointUpdateCallback)
        b->point.insert(b->point.end(), w, w + 3); b->point_keys.push_back(sym('m', cantor((uint64_t)o.tracklet, (uint64_t)o.frame)));
        ptp.idx.push_back(cam(o.frame)); ptp.idx.push_back(pi); ptp.meas.insert(ptp.meas.end(), o.z, o.z + 3);
        if (k > 0) { mot.idx.push_back(pi - 1); mot.idx.push_back(pi);
          if (wcpe) mot.idx.push_back(var_index[{o.object, dyn[t[k-1]].frame}]);
          mot.idx.push_back(var_index[{o.object, o.frame}]); }
      }
    }
    if (ptp.n()) b->blocks.push_back(std::move(ptp));
    if (mot.n()) b->blocks.push_back(std::move(mot));
    Block sm; sm.type = wcpe ? DYNOBA_SMOOTH_POSE6 : DYNOBA_BETWEEN6; sm.sigma.assign(P.smoothing_sigma, P.smoothing_sigma + 6); sm.sigma_dim = 6;
    const Pose I = identity_pose(); double pi12[12]; pose_to(I, pi12);
    for (auto& kv : var_index) {
      const int32_t obj = kv.first.first, fr = kv.first.second;
      auto p1 = var_index.find({obj, fr - 1}), p2 = var_index.find({obj, fr - 2});
      if (!wcpe) { if (p1 != var_index.end()) { sm.idx.push_back(p1->second); sm.idx.push_back(kv.second); sm.meas.insert(sm.meas.end(), pi12, pi12 + 12); } }
      else if (p1 != var_index.end() && p2 != var_index.end()) { sm.idx.push_back(p2->second); sm.idx.push_back(p1->second); sm.idx.push_back(kv.second); }
    }
    if (sm.n()) b->blocks.push_back(std::move(sm));
  } else {
  // ---- dynamic: objects -> visibility segments (key-frames) -> motion variables
  const int32_t n_cam = (int32_t)cam_index.size();
  std::map<int32_t, std::vector<int32_t>> obj_frames;                  // object -> sorted frames it is observed in
  for (auto& o : dyn) { BARG(cam(o.frame) >= 0, "dynamic observation in an unknown frame"); obj_frames[o.object].push_back(o.frame); }
  struct Seg { int32_t object, keyframe, aux; };
  std::map<std::pair<int32_t, int32_t>, int32_t> motion_index;         // (object, frame) -> pose index
  std::map<std::pair<int32_t, int32_t>, int32_t> seg_of;               // (object, frame) -> segment
  std::vector<Seg> segs; std::vector<std::vector<int32_t>> seg_motions;
  for (auto& kv : obj_frames) {
    auto& fr = kv.second; std::sort(fr.begin(), fr.end()); fr.erase(std::unique(fr.begin(), fr.end()), fr.end());
    for (size_t i = 0; i < fr.size(); i++) {
      if (i == 0 || fr[i] - fr[i-1] > P.keyframe_gap) { segs.push_back(Seg{ kv.first, fr[i], -1 }); seg_motions.emplace_back(); }
      const int32_t pi = n_cam + (int32_t)motion_index.size();
      motion_index[{kv.first, fr[i]}] = pi; seg_of[{kv.first, fr[i]}] = (int32_t)segs.size() - 1; seg_motions.back().push_back(pi);
      Pose H = identity_pose();                                        // key-frame motion starts at its prior; others at the front end's estimate
      if (fr[i] != segs.back().keyframe) { auto it = b->motion_init.find({kv.first, fr[i]}); if (it != b->motion_init.end()) H = it->second; }
      double p[12]; pose_to(H, p); b->pose.insert(b->pose.end(), p, p + 12);
      b->order.push_back(fr[i]); b->pose_keys.push_back(labeled('H', (uint64_t)('0' + kv.first), (uint64_t)fr[i]));
    }
  }
  // ---- dynamic tracklets (object-major, then first appearance), key-frame poses, hybrid factors
  {
    std::map<std::pair<int32_t, int64_t>, std::vector<size_t>> tracks; std::vector<std::pair<int32_t, int64_t>> first_seen;
    for (size_t i = 0; i < dyn.size(); i++) { auto key = std::make_pair(dyn[i].object, dyn[i].tracklet);
      auto& t = tracks[key]; if (t.empty()) first_seen.push_back(key); t.push_back(i); }
    std::stable_sort(first_seen.begin(), first_seen.end(), [](const std::pair<int32_t, int64_t>& a, const std::pair<int32_t, int64_t>& c) { return a.first < c.first; });
    // L_e per segment: given, or the centroid of the key-frame's world points with identity rotation
    for (size_t s = 0; s < segs.size(); s++) {
      Pose Le; auto it = b->keyframe_pose.find({segs[s].object, segs[s].keyframe});
      if (it != b->keyframe_pose.end()) Le = it->second;
      else {
        Le = identity_pose(); double c[3] = {0, 0, 0}; int cnt = 0;
        for (auto& o : dyn) if (o.object == segs[s].object && o.frame == segs[s].keyframe) { double w[3]; se3_transform_from(b->frames[o.frame].first, o.z, w); for (int k = 0; k < 3; k++) c[k] += w[k]; cnt++; }
        for (int k = 0; k < 3; k++) Le.t[k] = cnt ? c[k]/cnt : 0.0;
      }
      segs[s].aux = (int32_t)(b->aux.size()/12); double p[12]; pose_to(Le, p); b->aux.insert(b->aux.end(), p, p + 12);
    }
    Block blk; blk.type = DYNOBA_HYBRID3; blk.sigma = { P.sigma_dynamic }; blk.k = P.huber_k; blk.has_aux = true;
    for (auto& key : first_seen) {
      auto& t = tracks[key];
      const Obs& o0 = dyn[t[0]];
      const int32_t s0 = seg_of[{o0.object, o0.frame}];
      // every observation of a tracklet refers to the key-frame of its FIRST observation (a tracklet does not outlive a segment)
      bool one_segment = true; for (size_t i : t) one_segment = one_segment && seg_of[{dyn[i].object, dyn[i].frame}] == s0;
      BARG(one_segment, "a dynamic tracklet spans two key-frame segments of its object");
      const int32_t pi = (int32_t)(b->point.size()/3);
      // m_L = L_e^-1 (e_H_k)^-1 X_k z at the first observation (HybridObjectMotion::projectToObject3)
      const Pose Le = pose_from(&b->aux[(size_t)12*segs[s0].aux]); const Pose E = pose_from(&b->pose[(size_t)12*motion_index[{o0.object, o0.frame}]]);
      double w[3], q[3], m[3]; se3_transform_from(b->frames[o0.frame].first, o0.z, w); se3_transform_to(E, w, q); se3_transform_to(Le, q, m);
      b->point.insert(b->point.end(), m, m + 3); b->point_keys.push_back(sym('m', cantor((uint64_t)key.second, 0)));
      for (size_t i : t) { const Obs& o = dyn[i];
        blk.idx.push_back(cam(o.frame)); blk.idx.push_back(motion_index[{o.object, o.frame}]); blk.idx.push_back(pi);
        blk.meas.insert(blk.meas.end(), o.z, o.z + 3); blk.aux.push_back(segs[s0].aux); }
    }
    if (blk.n()) b->blocks.push_back(std::move(blk));
    // priors on the key-frame motions, three-motion smoothing inside a segment
    Block pr; pr.type = DYNOBA_PRIOR6; pr.sigma.assign(6, P.prior_sigma); pr.sigma_dim = 6;
    Block sm; sm.type = DYNOBA_SMOOTH_HYBRID6; sm.sigma.assign(P.smoothing_sigma, P.smoothing_sigma + 6); sm.sigma_dim = 6; sm.has_aux = true;
    const Pose I = identity_pose(); double pi12[12]; pose_to(I, pi12);
    for (size_t s = 0; s < segs.size(); s++) {
      pr.idx.push_back(seg_motions[s][0]); pr.meas.insert(pr.meas.end(), pi12, pi12 + 12);
      for (size_t i = 0; i + 2 < seg_motions[s].size(); i++) { sm.idx.push_back(seg_motions[s][i]); sm.idx.push_back(seg_motions[s][i + 1]); sm.idx.push_back(seg_motions[s][i + 2]); sm.aux.push_back(segs[s].aux); }
    }
    if (pr.n()) b->blocks.push_back(std::move(pr));
    if (sm.n()) b->blocks.push_back(std::move(sm));
  }
  }
  // ---- camera chain
  {
    Block pr; pr.type = DYNOBA_PRIOR6; pr.sigma.assign(6, P.prior_sigma); pr.sigma_dim = 6;
    pr.idx.push_back(0); pr.meas.insert(pr.meas.end(), b->pose.begin(), b->pose.begin() + 12);      // prior = the first frame's pose
    b->blocks.push_back(std::move(pr));
    Block od; od.type = DYNOBA_BETWEEN6; od.sigma.assign(P.odometry_sigma, P.odometry_sigma + 6); od.sigma_dim = 6;
    int32_t prev = -1;
    for (auto& kv : b->frames) {
      if (prev >= 0 && kv.second.second) { od.idx.push_back(cam(prev)); od.idx.push_back(cam(kv.first)); double p[12]; pose_to(b->odom[kv.first], p); od.meas.insert(od.meas.end(), p, p + 12); }
      prev = kv.first;
    }
    if (od.n()) b->blocks.push_back(std::move(od));
  }
  b->finalized = true;
  return DYNOBA_OK;
}

int dynoba_builder_counts(dynoba_builder_handle b, int64_t* n_pose, int64_t* n_point, int64_t* n_aux, int32_t* n_blocks) {
  BARG(b, "null builder"); int rc = dynoba_builder_finalize(b); if (rc) return rc;
  if (n_pose) *n_pose = (int64_t)b->pose.size()/12; if (n_point) *n_point = (int64_t)b->point.size()/3; if (n_aux) *n_aux = (int64_t)b->aux.size()/12;
  if (n_blocks) *n_blocks = (int32_t)b->blocks.size();
  return DYNOBA_OK;
}
int dynoba_builder_get_variables(dynoba_builder_handle b, double* pose, double* point, double* aux, int32_t* pose_order, uint64_t* pose_keys, uint64_t* point_keys) {
  BARG(b, "null builder"); int rc = dynoba_builder_finalize(b); if (rc) return rc;
  if (pose) std::copy(b->pose.begin(), b->pose.end(), pose); if (point) std::copy(b->point.begin(), b->point.end(), point);
  if (aux) std::copy(b->aux.begin(), b->aux.end(), aux); if (pose_order) std::copy(b->order.begin(), b->order.end(), pose_order);
  if (pose_keys) std::copy(b->pose_keys.begin(), b->pose_keys.end(), pose_keys); if (point_keys) std::copy(b->point_keys.begin(), b->point_keys.end(), point_keys);
  return DYNOBA_OK;
}
int dynoba_builder_block_info(dynoba_builder_handle b, int32_t bi, int32_t* type, int64_t* n, int32_t* sigma_dim, int64_t* sigma_count, double* robust_k, int32_t* has_aux) {
  BARG(b, "null builder"); int rc = dynoba_builder_finalize(b); if (rc) return rc;
  BARG(bi >= 0 && bi < (int32_t)b->blocks.size(), "bad block index");
  const Block& k = b->blocks[bi];
  if (type) *type = k.type; if (n) *n = k.n(); if (sigma_dim) *sigma_dim = k.sigma_dim; if (sigma_count) *sigma_count = 1;
  if (robust_k) *robust_k = k.k; if (has_aux) *has_aux = k.has_aux ? 1 : 0;
  return DYNOBA_OK;
}
int dynoba_builder_get_block(dynoba_builder_handle b, int32_t bi, int32_t* idx, double* meas, double* sigma, int32_t* aux) {
  BARG(b, "null builder"); int rc = dynoba_builder_finalize(b); if (rc) return rc;
  BARG(bi >= 0 && bi < (int32_t)b->blocks.size(), "bad block index");
  const Block& k = b->blocks[bi];
  if (idx) std::copy(k.idx.begin(), k.idx.end(), idx); if (meas) std::copy(k.meas.begin(), k.meas.end(), meas);
  if (sigma) std::copy(k.sigma.begin(), k.sigma.end(), sigma); if (aux) std::copy(k.aux.begin(), k.aux.end(), aux);
  return DYNOBA_OK;
}
// hands everything to a solver handle: the calls a flattening adapter would make, without the factor objects in between
int dynoba_builder_emit(dynoba_builder_handle b, dynoba_handle h) {
  BARG(b && h, "null"); int rc = dynoba_builder_finalize(b); if (rc) return rc;
  if ((rc = dynoba_set_variables(h, DYNOBA_POSE6, (int64_t)b->pose.size()/12, b->pose_keys.data(), b->pose.data()))) return rc;
  if ((rc = dynoba_set_variables(h, DYNOBA_POINT3, (int64_t)b->point.size()/3, b->point_keys.data(), b->point.data()))) return rc;
  if (!b->aux.empty() && (rc = dynoba_set_aux_poses(h, (int64_t)b->aux.size()/12, b->aux.data()))) return rc;
  for (auto& k : b->blocks)
    if ((rc = dynoba_add_factors(h, k.type, k.n(), k.idx.data(), k.meas.empty() ? nullptr : k.meas.data(), k.sigma.data(), k.sigma_dim, 1, k.k, k.has_aux ? k.aux.data() : nullptr))) return rc;
  return dynoba_set_pose_order(h, (int64_t)b->order.size(), b->order.data());
}

}  // extern "C"
