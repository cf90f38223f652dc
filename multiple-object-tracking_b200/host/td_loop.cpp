// td_loop.cpp -- the per-frame tracking control flow of the reference, kept source-compatible but batched.
//
// Mirrors one iteration of pthread_mtcnn_trkn (top/td.cpp:343-644) for one or several independent streams:
//   predict every live track (+ clamp, :344-384)  ->  cost matrix + assignmentoptimal (:386-470)  ->  scatter (:472-502)
//   ->  update assigned tracks with their detection (:512-547)  ->  age + update unassigned tracks with their own box
//   (:550-582)  ->  delete lost tracks with stable compaction (:585-609)  ->  spawn a tracker per unassigned detection
//   in ascending order (:612-644; with KCF_TRACKER the new tracker gets its first update at once).
// Where the reference loops over tracks and calls the plugin once per track, this calls the batched C ABI once per
// stage for ALL tracks of ALL streams stepped together; the bookkeeping (ages, visibility counters, ids, order of the
// track table) is the reference's, per stream.  Only the C ABI of include/mot_b200.h is used.
#include "../../include/mot_b200.h"

#include <cstring>
#include <vector>

namespace {

struct TrackInfo {                       // tracker_info_t, top/td.cpp:271-290 (without the drawing colour and the scratch image)
    uint32_t tid;
    int handle;
    int age, totalVisibleCount, consecutiveInvisibleCount;
    mot_bbox_t bbox;
};

constexpr int MAX_INVISIBLE_COUNTS = 20;  // top/td.cpp:267
constexpr int MIN_AGE_COUNTS = 10;        // top/td.cpp:268
// VISIBILITY_THRESHOLD 0.60 is applied as visible*5 < 3*age (top/td.cpp:589)

}  // namespace

struct mot_td_s {
    mot_ctx_t *ctx;
    int frame_slot, cap, cost_mode, kcf;
    uint32_t tracker_id = 0;
    long dropped = 0;                    // detections that could not spawn a track (window the KCF plugin cannot serve)
    std::vector<TrackInfo> tracks;
    std::vector<mot_bbox_t> predicted;
    std::vector<int> assigned_trackers, assigned_detected;
};

extern "C" {

int mot_td_create(mot_td_t **out, mot_ctx_t *ctx, int frame_slot, int cap, int cost_mode)
{
    if (!out || !ctx || cap <= 0) return MOT_ERR_ARG;
    mot_td_t *td = new mot_td_s();
    td->ctx = ctx; td->frame_slot = frame_slot; td->cap = cap; td->cost_mode = cost_mode;
    td->kcf = mot_ctx_kind(ctx) == MOT_TRACKER_KCF;
    *out = td;
    return 0;
}

void mot_td_destroy(mot_td_t *td)
{
    if (!td) return;
    std::vector<int> h;
    for (auto &t : td->tracks) h.push_back(t.handle);
    if (!h.empty()) mot_tracker_delete_batch(td->ctx, (int)h.size(), h.data());
    delete td;
}

int mot_td_step_multi(mot_td_t **tds, int ns, const uint8_t *const *host_bgr, int stride_bytes, const mot_bbox_t *const *dets, const int *ndet)
{
    if (!tds || ns <= 0 || !dets || !ndet) return MOT_ERR_ARG;
    mot_ctx_t *ctx = tds[0]->ctx;
    int rc;
    for (int s = 0; s < ns; ++s) {
        if (tds[s]->ctx != ctx) return MOT_ERR_ARG;
        if (host_bgr && host_bgr[s]) { rc = mot_frame_upload(ctx, tds[s]->frame_slot, host_bgr[s], stride_bytes); if (rc) return rc; }
    }

    // ---- predict all + clamp (top/td.cpp:344-384) --------------------------------------------------------------
    std::vector<int> handles, frames; std::vector<mot_bbox_t> boxes;
    for (int s = 0; s < ns; ++s)
        for (auto &t : tds[s]->tracks) { handles.push_back(t.handle); frames.push_back(tds[s]->frame_slot); boxes.push_back(t.bbox); }
    rc = mot_predict_batch(ctx, (int)handles.size(), handles.data(), frames.data(), boxes.data(), 1);
    if (rc) return rc;
    {
        size_t k = 0;
        for (int s = 0; s < ns; ++s) {
            tds[s]->predicted.clear();
            for (auto &t : tds[s]->tracks) { t.bbox = boxes[k++]; tds[s]->predicted.push_back(t.bbox); }
        }
    }

    // ---- cost matrices + assignment (top/td.cpp:386-470) ----------------------------------------------------------
    {
        std::vector<int> T(ns), D(ns);
        long mt = 1, mdd = 1;
        for (int s = 0; s < ns; ++s) { T[s] = (int)tds[s]->tracks.size(); D[s] = ndet[s]; if (T[s] > mt) mt = T[s]; if (D[s] > mdd) mdd = D[s]; }
        std::vector<mot_bbox_t> trk((size_t)mt * ns), det((size_t)mdd * ns);
        const long md = mt > mdd ? mt : mdd;
        std::vector<int> assign((size_t)md * ns, -1);
        for (int s = 0; s < ns; ++s) {
            for (int i = 0; i < T[s]; ++i) trk[(size_t)mt * s + i] = tds[s]->tracks[i].bbox;
            if (D[s]) memcpy(&det[(size_t)mdd * s], dets[s], sizeof(mot_bbox_t) * D[s]);
        }
        rc = mot_associate_batch(ctx, ns, T.data(), D.data(), trk.data(), mt, det.data(), mdd, tds[0]->cost_mode,
                                 nullptr, 0, assign.data(), md, nullptr);
        if (rc) return rc;
        // scatter (top/td.cpp:472-502)
        for (int s = 0; s < ns; ++s) {
            mot_td_t *td = tds[s];
            td->assigned_trackers.assign(T[s], -1); td->assigned_detected.assign(D[s], -1);
            if (!T[s] || !D[s]) continue;                      // assignmentoptimal is not called (:460)
            const int *a = &assign[(size_t)md * s];
            if (T[s] < D[s]) { for (int i = 0; i < T[s]; ++i) { const int j = a[i]; td->assigned_trackers[i] = j; if (j >= 0) td->assigned_detected[j] = i; } }
            else             { for (int j = 0; j < D[s]; ++j) { const int i = a[j]; if (i >= 0) td->assigned_trackers[i] = j; td->assigned_detected[j] = i; } }
        }
    }

    // ---- update assigned (:512-547) and unassigned (:550-582) ------------------------------------------------------
    handles.clear(); frames.clear(); boxes.clear();
    for (int s = 0; s < ns; ++s) {
        mot_td_t *td = tds[s];
        for (size_t i = 0; i < td->tracks.size(); ++i) {
            TrackInfo &t = td->tracks[i];
            const int j = td->assigned_trackers[i];
            if (j >= 0) { t.bbox = dets[s][j]; t.totalVisibleCount++; t.age++; t.consecutiveInvisibleCount = 0; }
            else        { t.age++; t.consecutiveInvisibleCount++; }
            handles.push_back(t.handle); frames.push_back(td->frame_slot); boxes.push_back(t.bbox);
        }
    }
    rc = mot_update_batch(ctx, (int)handles.size(), handles.data(), frames.data(), boxes.data());
    if (rc) return rc;

    // ---- delete lost (:585-609) ----------------------------------------------------------------------------------------
    std::vector<int> dead;
    for (int s = 0; s < ns; ++s) {
        mot_td_t *td = tds[s];
        size_t n = 0;
        for (size_t i = 0; i < td->tracks.size(); ++i) {
            const TrackInfo &t = td->tracks[i];
            const bool lost = ((t.age < MIN_AGE_COUNTS) && (t.totalVisibleCount * 5 < 3 * t.age)) || (t.consecutiveInvisibleCount >= MAX_INVISIBLE_COUNTS);
            if (!lost) { if (n != i) td->tracks[n] = td->tracks[i]; ++n; }
            else dead.push_back(t.handle);
        }
        td->tracks.resize(n);
    }
    if (!dead.empty()) { rc = mot_tracker_delete_batch(ctx, (int)dead.size(), dead.data()); if (rc) return rc; }

    // ---- spawn (:612-644) ------------------------------------------------------------------------------------------------
    std::vector<mot_bbox_t> nb; std::vector<int> nframe; std::vector<std::pair<int, size_t>> where;
    for (int s = 0; s < ns; ++s) {
        mot_td_t *td = tds[s];
        for (int j = 0; j < ndet[s]; ++j) {
            if (td->assigned_detected[j] >= 0) continue;
            if ((int)td->tracks.size() >= td->cap) break;       // the reference has no guard (256-slot stack array, top/td.cpp:312)
            // a detection the plugin cannot build a tracker for (KCF: smaller than 2x2 cells or larger than the frame) is counted
            // and skipped BEFORE anything is mutated, like the device-resident loop does: one bad box must not poison the table
            if (!mot_tracker_spawnable(ctx, &dets[s][j])) { td->dropped++; continue; }
            TrackInfo t{};
            t.tid = td->tracker_id++; t.bbox = dets[s][j]; t.handle = -1;
            td->tracks.push_back(t);
            nb.push_back(dets[s][j]); nframe.push_back(td->frame_slot); where.push_back({ s, td->tracks.size() - 1 });
        }
    }
    if (!nb.empty()) {
        std::vector<int> nh(nb.size());
        rc = mot_tracker_new_batch(ctx, (int)nb.size(), nb.data(), nh.data());
        if (rc) {
            // nothing was created (the call validates before it allocates): take the half-made entries out again so that the
            // next step sees a consistent table; the ids are handed back too
            for (size_t i = where.size(); i-- > 0;) {
                mot_td_t *td = tds[where[i].first];
                td->tracks.pop_back(); td->tracker_id--;
            }
            return rc;
        }
        for (size_t i = 0; i < nb.size(); ++i) tds[where[i].first]->tracks[where[i].second].handle = nh[i];
        if (tds[0]->kcf) { rc = mot_update_batch(ctx, (int)nb.size(), nh.data(), nframe.data(), nb.data()); if (rc) return rc; }   // first update (:629-641)
    }
    return 0;
}

int mot_td_step(mot_td_t *td, const uint8_t *host_bgr, int stride_bytes, const mot_bbox_t *dets, int ndet)
{
    const uint8_t *f[1] = { host_bgr };
    const mot_bbox_t *d[1] = { dets };
    return mot_td_step_multi(&td, 1, f, stride_bytes, d, &ndet);
}

// ---- the detector's wire format ----------------------------------------------------------------------------------
// The tracking thread receives one bbox_chain_t per frame (top/td.cpp:326-335 reads pdetected->nbox and pdetected->bbox[]); the detector
// thread fills up to MAX_GPU_BATCH = 4 of them per tensorRunB call (top/td.cpp:178-204) and queues them in order.
int mot_td_step_chain(mot_td_t *td, const uint8_t *host_bgr, int stride_bytes, const mot_bbox_chain_t *chain)
{
    if (!td || !chain || chain->nbox < 0 || chain->nbox > 128) return MOT_ERR_ARG;
    return mot_td_step(td, host_bgr, stride_bytes, chain->bbox, chain->nbox);
}

int mot_td_step_chain_batch(mot_td_t *td, int nbatch, const uint8_t *const *host_bgr, int stride_bytes, const mot_bbox_chain_t *const *chains)
{
    if (!td || nbatch < 0 || nbatch > 4 || (nbatch && !chains)) return MOT_ERR_ARG;
    for (int i = 0; i < nbatch; ++i) {                       // consumed in queue order, one frame at a time, like prc_rptr walks the ring
        const int rc = mot_td_step_chain(td, host_bgr ? host_bgr[i] : nullptr, stride_bytes, chains[i]);
        if (rc) return rc;
    }
    return 0;
}

int mot_td_ntracks(mot_td_t *td) { return td ? (int)td->tracks.size() : 0; }
long mot_td_dropped(mot_td_t *td) { return td ? td->dropped : 0; }

void mot_td_get(mot_td_t *td, uint32_t *tid, mot_bbox_t *boxes, int *age, int *vis, int *invis)
{
    for (size_t i = 0; i < td->tracks.size(); ++i) {
        const TrackInfo &t = td->tracks[i];
        if (tid) tid[i] = t.tid;
        if (boxes) boxes[i] = t.bbox;
        if (age) age[i] = t.age;
        if (vis) vis[i] = t.totalVisibleCount;
        if (invis) invis[i] = t.consecutiveInvisibleCount;
    }
}

// The overlay of top/td.cpp:647-733 for the current track table: three nested rectangles per track, in table order, in the
// colour colormap[hashcolor(tid + 1) & 255] that the reference fixes at spawn time (top/td.cpp:619-620), drawn on the device into
// the loop's frame slot.  mot_frame_download brings the annotated frame back.
int mot_td_overlay(mot_td_t *td)
{
    if (!td) return MOT_ERR_ARG;
    const int n = (int)td->tracks.size();
    std::vector<int> slots(n, td->frame_slot);
    std::vector<mot_bbox_t> boxes(n);
    std::vector<uint32_t> rgb(n);
    for (int i = 0; i < n; ++i) { boxes[i] = td->tracks[i].bbox; rgb[i] = mot_track_color(td->tracks[i].tid); }
    return mot_overlay_batch(td->ctx, n, slots.data(), boxes.data(), rgb.data(), 3);
}

int mot_td_last(mot_td_t *td, mot_bbox_t *predicted, int *assigned_trackers)
{
    const size_t n = td->predicted.size();
    for (size_t i = 0; i < n; ++i) {
        if (predicted) predicted[i] = td->predicted[i];
        if (assigned_trackers) assigned_trackers[i] = td->assigned_trackers[i];
    }
    return (int)n;
}

}  // extern "C"
