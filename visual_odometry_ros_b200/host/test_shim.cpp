// test_shim.cpp -- exercises the reference-API shim classes end to end on a GPU (run by tests/test_host_shim.py).
#include "vo_shim.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#define CHECK(cond, msg) do { if (!(cond)) { printf("SHIM_FAIL %s:%d %s\n", __FILE__, __LINE__, msg); return 1; } } while (0)

static unsigned lcg_state = 12345u;
static float frand() { lcg_state = lcg_state * 1664525u + 1013904223u; return (float)((lcg_state >> 8) & 0xffffff) / 16777216.f; }

// smooth random texture, analytically shiftable
static float tex(float x, float y)
{
    float v = 0.f;
    const float fx[6] = {0.11f, 0.23f, 0.37f, 0.05f, 0.17f, 0.29f}, fy[6] = {0.13f, 0.07f, 0.31f, 0.19f, 0.41f, 0.03f};
    const float ph[6] = {0.3f, 1.1f, 2.0f, 4.2f, 5.5f, 0.7f};
    for (int k = 0; k < 6; ++k) v += std::sin(fx[k] * x + fy[k] * y + ph[k]) * std::cos(fy[k] * x - fx[k] * y + 2 * ph[k]);
    return 128.f + 20.f * v;
}
static void render(std::vector<unsigned char> &buf, int w, int h, float dx, float dy)
{
    buf.resize((size_t)w * h);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        float v = tex(x - dx, y - dy);
        buf[(size_t)y * w + x] = (unsigned char)std::fmin(255.f, std::fmax(0.f, std::floor(v + 0.5f)));
    }
}

template <class F> static bool throws_with(F f, const char *needle)
{
    try { f(); } catch (const std::runtime_error &e) { return std::string(e.what()).find(needle) != std::string::npos; }
    return false;
}

int main()
{
    const int W = 640, H = 200;
    std::vector<unsigned char> b0, b1;
    const float sx = 2.4f, sy = -1.3f;
    render(b0, W, H, 0, 0);
    render(b1, W, H, sx, sy);
    cv::Mat img0(H, W, b0.data()), img1(H, W, b1.data());
    PixelVec pts0;
    for (int y = 30; y < H - 30; y += 14) for (int x = 30; x < W - 30; x += 20) pts0.emplace_back(x + frand(), y + frand());
    const size_t n = pts0.size();

    // ---- FeatureTracker
    FeatureTracker ft;
    PixelVec p1; MaskVec m;
    ft.track(img0, img1, pts0, 21, 3, 30.f, p1, m);
    CHECK(p1.size() == n && m.size() == n, "track sizes");
    size_t good = 0;
    for (size_t i = 0; i < n; ++i) if (m[i] && std::fabs(p1[i].x - pts0[i].x - sx) < 0.1f && std::fabs(p1[i].y - pts0[i].y - sy) < 0.1f) ++good;
    CHECK(good > n * 9 / 10, "track accuracy");
    PixelVec p2 = pts0; MaskVec m2(n, true); m2[3] = false;
    ft.trackWithPrior(img0, img1, pts0, 21, 3, 30.f, p2, m2);
    CHECK(!m2[3], "pre-existing false mask entries must survive (Appendix B #7)");
    PixelVec p3; MaskVec m3;
    ft.trackBidirection(img0, img1, pts0, 21, 3, 30.f, 0.5f, p3, m3);
    PixelVec p4 = p1; MaskVec m4;
    ft.trackBidirectionWithPrior(img0, img1, pts0, 21, 3, 30.f, 0.5f, p4, m4);
    size_t c3 = 0, c4 = 0;
    for (size_t i = 0; i < n; ++i) { c3 += m3[i]; c4 += m4[i]; }
    CHECK(c3 > n * 8 / 10 && c4 > n * 8 / 10, "bidirectional survivors");
    std::vector<float> scale(n, 1.0f);
    PixelVec p5 = p1; MaskVec m5;
    ft.trackWithScale(img0, cv::Mat(), cv::Mat(), img1, pts0, scale, p5, m5);
    size_t c5 = 0;
    for (size_t i = 0; i < n; ++i) c5 += (m5[i] && std::fabs(p5[i].x - pts0[i].x - sx) < 0.15f);
    CHECK(c5 > n * 8 / 10, "trackWithScale accuracy");
    PixelVec bad(3);
    CHECK(throws_with([&] { MaskVec mm; ft.trackWithScale(img0, cv::Mat(), cv::Mat(), img1, pts0, scale, bad, mm); }, "pts_track.size() != pts0.size()"),
          "trackWithScale size error text");

    // ---- FeatureExtractor (reference API; cv::ORB keypoints restated on the device)
    {
        FeatureExtractor fe;
        fe.initParams(img0.cols, img0.rows, 16, 8, 20, 5);
        fe.resetWeightBin();
        PixelVec e0, e1;
        fe.extractORBwithBinning_fast(img0, e0, true);
        CHECK(e0.size() > 20 && e0.size() <= 16 * 8, "bucketed ORB extraction");
        fe.updateWeightBin(e0);
        fe.extractORBwithBinning_fast(img0, e1, true);
        CHECK(e1.size() < e0.size(), "occupied buckets are skipped");
        FeatureExtractor fe2;
        CHECK(throws_with([&] { PixelVec q; fe2.extractORBwithBinning_fast(img0, q, true); }, "initParams"), "extractor without initParams");
    }

    // ---- MotionEstimator
    const float fx = 718.856f, fy = 718.856f, cx = 607.19f, cy = 185.21f, base = 0.537f;
    PointVec X; PixelVec pl, pr;
    const float tz = 0.8f, tx = 0.03f;
    for (int i = 0; i < 400; ++i) {
        Point P; P(0) = -10.f + 20.f * frand(); P(1) = -3.f + 5.f * frand(); P(2) = 5.f + 40.f * frand();
        X.push_back(P);
        const float xl = P(0) - tx, yl = P(1), zl = P(2) - tz;      // T10 = translate(-t01)
        pl.emplace_back(fx * xl / zl + cx + 0.2f * (frand() - 0.5f), fy * yl / zl + cy + 0.2f * (frand() - 0.5f));
        pr.emplace_back(fx * (xl - base) / zl + cx + 0.2f * (frand() - 0.5f), fy * yl / zl + cy + 0.2f * (frand() - 0.5f));
    }
    PoseSE3 Tlr = PoseSE3::Identity(); Tlr(0, 3) = base;
    MotionEstimator me(true, Tlr);
    PoseSE3 T01 = PoseSE3::Identity(); MaskVec mi;
    auto cam = std::make_shared<Camera>(fx, fy, cx, cy);
    CameraConstPtr camc = cam;
    CHECK(me.poseOnlyBundleAdjustment_Stereo(X, pl, pr, camc, camc, Tlr, 3.0f, T01, mi), "stereo pose success");
    CHECK(std::fabs(T01(2, 3) - tz) < 5e-3f && std::fabs(T01(0, 3) - tx) < 5e-3f, "stereo pose value");
    PoseSE3 T01b = PoseSE3::Identity(); MaskVec mib;
    me.poseOnlyBundleAdjustment_Stereo(X, pl, pr, fx, fy, cx, cy, fx, fy, cx, cy, Tlr, 3.0f, T01b, mib);
    CHECK(std::fabs(T01b(2, 3) - T01(2, 3)) < 1e-6f, "standalone overload agrees");
    Rot3 R = Rot3::Identity(); Pos3 t; MaskVec mm;
    CHECK(me.poseOnlyBundleAdjustment(X, pl, camc, 5, R, t, mm), "mono pose success");
    CHECK(std::fabs(t(2) - tz) < 2e-2f, "mono pose value");
    MotionEstimator me_mono(false);
    CHECK(throws_with([&] { PoseSE3 T = PoseSE3::Identity(); MaskVec q; me_mono.poseOnlyBundleAdjustment_Stereo(X, pl, pr, camc, camc, Tlr, 3.f, T, q); },
                      "is_stereo_mode_ == false"), "mode error text");
    PixelVec shortv(5);
    CHECK(throws_with([&] { PoseSE3 T = PoseSE3::Identity(); MaskVec q; me.poseOnlyBundleAdjustment_Stereo(X, shortv, pr, camc, camc, Tlr, 3.f, T, q); },
                      "X.size() != pts_l1.size()"), "size error text");

    // ---- mono geometric front-end: five-point, Sampson / symmetric epipolar distance, 1-point voting
    {
        PixelVec m0, m1;
        const float yaw = 0.02f, cyw = std::cos(yaw), syw = std::sin(yaw);
        const float tx = 0.9f * std::sin(0.5f * yaw), tzz = 0.9f * std::cos(0.5f * yaw);               // planar motion: X1 = Ry(yaw) X0 + t
        for (size_t i = 0; i < X.size(); ++i) {
            const float x1 = cyw * X[i](0) + syw * X[i](2) + tx, y1 = X[i](1), z1 = -syw * X[i](0) + cyw * X[i](2) + tzz;
            m0.emplace_back(fx * X[i](0) / X[i](2) + cx, fy * X[i](1) / X[i](2) + cy);
            m1.emplace_back(fx * x1 / z1 + cx, fy * y1 / z1 + cy);
        }
        MotionEstimator mm5(false);
        mm5.setThres5p(1.0f);
        Rot3 R5; Pos3 t5; PointVec X5; MaskVec k5;
        CHECK(mm5.calcPose5PointsAlgorithm(m0, m1, camc, R5, t5, X5, k5), "five-point success");
        const float tn = std::sqrt(tx * tx + tzz * tzz);
        CHECK(std::fabs(R5(0, 2) - syw) < 2e-3f && std::fabs(t5(2) - tzz / tn) < 2e-2f, "five-point pose value");
        size_t n_in = 0; for (size_t i = 0; i < k5.size(); ++i) n_in += k5[i];
        CHECK(n_in > k5.size() * 9 / 10 && X5.size() == m0.size(), "five-point inliers");
        std::vector<float> ds, de;
        mm5.calcSampsonDistance(m0, m1, camc, R5, t5, ds);
        mm5.calcSymmetricEpipolarDistance(m0, m1, camc, R5, t5, de);
        float worst_s = 0, worst_e = 0;
        for (size_t i = 0; i < ds.size(); ++i) { worst_s = std::fmax(worst_s, ds[i]); worst_e = std::fmax(worst_e, de[i]); }
        CHECK(ds.size() == m0.size() && worst_s < 1.0f && worst_e < 2.0f, "epipolar distances of exact correspondences");
        mm5.setThres1p(5.0f);
        MaskVec k1;
        const float th = mm5.findInliers1PointHistogram(m0, m1, camc, k1);
        CHECK(std::fabs(th - yaw) < 0.01f || std::fabs(th + yaw) < 0.01f, "1-point histogram finds the yaw");
        PixelVec shortp(3);
        CHECK(throws_with([&] { mm5.calcPose5PointsAlgorithm(m0, shortp, camc, R5, t5, X5, k5); }, "pts0.size() != pts1.size()"), "five-point size error text");
        CHECK(throws_with([&] { mm5.calcSymmetricEpipolarDistance(m0, shortp, camc, R5, t5, de); }, "calcSymmetricEpipolarDistance"), "epipolar size error text");
    }

    // ---- triangulateDLT
    Rot3 R10 = Rot3::Identity(); Pos3 t10; t10(0) = -base;
    PixelVec q0, q1;
    for (int i = 0; i < 100; ++i) { q0.emplace_back(fx * X[i](0) / X[i](2) + cx, fy * X[i](1) / X[i](2) + cy); q1.emplace_back(fx * (X[i](0) - base) / X[i](2) + cx, fy * X[i](1) / X[i](2) + cy); }
    PointVec X0, X1;
    mapping::triangulateDLT(q0, q1, R10, t10, camc, X0, X1);
    double worst = 0;
    for (int i = 0; i < 100; ++i) worst = std::fmax(worst, std::fabs(X0[i](2) - X[i](2)) / X[i](2));
    CHECK(worst < 2e-3, "triangulation accuracy");
    Point a0, a1;
    mapping::triangulateDLT(q0[7], q1[7], R10, t10, camc, camc, a0, a1);
    CHECK(a0(0) == X0[7](0) && a0(2) == X0[7](2), "scalar and vector triangulation agree bit for bit");

    // ---- DepthFilter
    DepthFilter df;
    double xu, cu;
    df.updateNormalDistribution(0.2, 1e-3, 0.25, 2e-3, xu, cu);
    CHECK(std::fabs(cu - (1e-3 * 2e-3) / 3e-3) < 1e-18 && std::fabs(xu - (0.2 * 2e-3 + 0.25 * 1e-3) / 3e-3) < 1e-15, "depth filter");

    // ---- local BA through the reference-shaped classes
    const int NKF = 5;
    std::vector<FramePtr> lefts, rights;
    for (int k = 0; k < NKF; ++k) {
        auto l = std::make_shared<Frame>(false), r = std::make_shared<Frame>(true);
        PoseSE3 Twc = PoseSE3::Identity(); Twc(2, 3) = 1.0f * k; Twc(0, 3) = 0.01f * k;
        if (k >= 2) { Twc(2, 3) += 0.03f * (frand() - 0.5f); Twc(0, 3) += 0.03f * (frand() - 0.5f); }   // perturbed
        l->setPose(Twc);
        PoseSE3 Twr = Twc; Twr(0, 3) += base;
        r->setPose(Twr); r->setLeftFramePtr(l);
        lefts.push_back(l); rights.push_back(r);
    }
    std::vector<LandmarkPtr> lms;
    for (int i = 0; i < 120; ++i) {
        auto lm = std::make_shared<Landmark>();
        Point Xw; Xw(0) = -8.f + 16.f * frand(); Xw(1) = -2.f + 4.f * frand(); Xw(2) = 8.f + 30.f * frand();
        for (int k = 0; k < NKF; ++k) {
            const float xl = Xw(0) - 0.01f * k, yl = Xw(1), zl = Xw(2) - 1.0f * k;       // true poses
            lm->addObservationOnKeyframe(Pixel(fx * xl / zl + cx, fy * yl / zl + cy), lefts[k]);
            lm->addObservationOnKeyframe(Pixel(fx * (xl - base) / zl + cx, fy * yl / zl + cy), rights[k]);
            lefts[k]->addRelatedLandmark(lm);
        }
        Point Xn = Xw; Xn(2) *= 1.f + 0.03f * (frand() - 0.5f);
        lm->set3DPoint(Xn);
        lms.push_back(lm);
    }
    FramePtrVec frames_ba;
    for (auto &l : lefts) frames_ba.push_back(l);
    for (auto &r : rights) frames_ba.push_back(r);
    std::vector<int> idx_fix = {0, 1}, idx_opt;
    for (int j = 2; j < (int)frames_ba.size(); ++j) if (!frames_ba[j]->isRightImage()) idx_opt.push_back(j);   // motion_estimator.cpp:1285-1293
    auto params = std::make_shared<SparseBAParameters>(true, Tlr);
    params->setPosesAndPoints(frames_ba, idx_fix, idx_opt);
    CHECK(params->getNumOfOptimizeFrames() == NKF - 2 && params->getNumOfOptimizeLandmarks() == 120 && params->getNumOfObservations() == 120 * NKF * 2, "BA packing");
    SparseBundleAdjustmentSolver solver(true);
    solver.setStereoCameras(cam, cam);
    solver.setBAParameters(params);
    solver.setHuberThreshold(0.5);
    const PoseSE3 before = lefts[3]->getPose();
    solver.solveForFiniteIterations(10);
    CHECK(lms[0]->isBundled(), "landmarks bundled");
    CHECK(std::fabs(lefts[3]->getPose()(2, 3) - 3.0f) <= std::fabs(before(2, 3) - 3.0f) + 1e-3f, "BA did not diverge");
    CHECK(lefts[0]->getPose()(2, 3) == 0.0f, "fixed keyframe untouched");
    SparseBundleAdjustmentSolver mono_solver(false);
    CHECK(throws_with([&] { mono_solver.setStereoCameras(cam, cam); }, "'is_stereo' should be set to 'true'"), "BA mode error text");

    // ---- the drivers the VO classes call (motion_estimator.h:122-123): same window through
    //      MotionEstimator::localBundleAdjustmentSparseSolver_Stereo must give the same result as the explicit sequence above
    {
        std::vector<FramePtr> l2, r2;
        std::vector<LandmarkPtr> lm2;
        lcg_state = 777u;
        auto build = [&](std::vector<FramePtr> &L, std::vector<FramePtr> &R, std::vector<LandmarkPtr> &LM) {
            lcg_state = 777u;
            for (int k = 0; k < NKF; ++k) {
                auto l = std::make_shared<Frame>(false), r = std::make_shared<Frame>(true);
                PoseSE3 Twc = PoseSE3::Identity(); Twc(2, 3) = 1.0f * k; Twc(0, 3) = 0.01f * k;
                if (k >= 2) { Twc(2, 3) += 0.03f * (frand() - 0.5f); Twc(0, 3) += 0.03f * (frand() - 0.5f); }
                l->setPose(Twc);
                PoseSE3 Twr = Twc; Twr(0, 3) += base;
                r->setPose(Twr); r->setLeftFramePtr(l);
                L.push_back(l); R.push_back(r);
            }
            for (int i = 0; i < 90; ++i) {
                auto lm = std::make_shared<Landmark>();
                Point Xw; Xw(0) = -8.f + 16.f * frand(); Xw(1) = -2.f + 4.f * frand(); Xw(2) = 8.f + 30.f * frand();
                for (int k = 0; k < NKF; ++k) {
                    const float xl = Xw(0) - 0.01f * k, yl = Xw(1), zl = Xw(2) - 1.0f * k;
                    lm->addObservationOnKeyframe(Pixel(fx * xl / zl + cx, fy * yl / zl + cy), L[k]);
                    lm->addObservationOnKeyframe(Pixel(fx * (xl - base) / zl + cx, fy * yl / zl + cy), R[k]);
                    L[k]->addRelatedLandmark(lm);
                }
                Point Xn = Xw; Xn(2) *= 1.f + 0.03f * (frand() - 0.5f);
                lm->set3DPoint(Xn);
                LM.push_back(lm);
            }
        };
        std::vector<FramePtr> la, ra, lb, rb;
        std::vector<LandmarkPtr> lma, lmb;
        build(la, ra, lma);
        build(lb, rb, lmb);
        // (a) explicit sequence
        {
            FramePtrVec fb;
            for (auto &l : la) fb.push_back(l);
            for (auto &r : ra) fb.push_back(r);
            std::vector<int> fix = {0, 1}, opt;
            for (int j = 2; j < (int)fb.size(); ++j) if (!fb[j]->isRightImage()) opt.push_back(j);
            auto prm = std::make_shared<SparseBAParameters>(true, Tlr);
            prm->setPosesAndPoints(fb, fix, opt);
            SparseBundleAdjustmentSolver sv(true);
            sv.setStereoCameras(cam, cam); sv.setBAParameters(prm); sv.setHuberThreshold(0.5);
            sv.solveForFiniteIterations(10);
        }
        // (b) the driver over a StereoKeyframes window
        auto win = std::make_shared<StereoKeyframes>();
        win->setMaxStereoKeyframes(NKF);
        MotionEstimator me_st(true, Tlr);
        win->addNewStereoKeyframe(std::make_shared<StereoFrame>(lb[0], rb[0]));
        win->addNewStereoKeyframe(std::make_shared<StereoFrame>(lb[1], rb[1]));
        CHECK(!me_st.localBundleAdjustmentSparseSolver_Stereo(win, camc, camc, Tlr), "fewer than 3 keyframes: local BA skipped");
        for (int k = 2; k < NKF; ++k) win->addNewStereoKeyframe(std::make_shared<StereoFrame>(lb[k], rb[k]));
        CHECK(me_st.localBundleAdjustmentSparseSolver_Stereo(win, camc, camc, Tlr), "stereo local-BA driver runs");
        bool same = true;
        for (int k = 0; k < NKF; ++k) for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) same &= la[k]->getPose()(r, c) == lb[k]->getPose()(r, c);
        for (size_t i = 0; i < lma.size(); ++i) for (int r = 0; r < 3; ++r) same &= lma[i]->get3DPoint()(r) == lmb[i]->get3DPoint()(r);
        CHECK(same, "driver result == explicit SparseBAParameters + solver sequence");
        CHECK(lmb[0]->isBundled(), "driver bundles the landmarks");
        MotionEstimator me_mono(false);
        CHECK(throws_with([&] { me_mono.localBundleAdjustmentSparseSolver_Stereo(win, camc, camc, Tlr); }, "is_stereo_mode_ == false"),
              "stereo driver mode error text");
        // mono driver: left frames only, left observations only
        auto kwin = std::make_shared<Keyframes>();
        kwin->setMaxKeyframes(NKF);
        std::vector<FramePtr> lm_f;
        std::vector<LandmarkPtr> lm_l;
        lcg_state = 4242u;
        for (int k = 0; k < NKF; ++k) {
            auto l = std::make_shared<Frame>(false);
            PoseSE3 Twc = PoseSE3::Identity(); Twc(2, 3) = 1.0f * k; Twc(0, 3) = 0.3f * k;
            if (k >= 2) { Twc(2, 3) += 0.02f * (frand() - 0.5f); Twc(0, 3) += 0.02f * (frand() - 0.5f); }
            l->setPose(Twc);
            lm_f.push_back(l);
        }
        for (int i = 0; i < 90; ++i) {
            auto lm = std::make_shared<Landmark>();
            Point Xw; Xw(0) = -8.f + 16.f * frand(); Xw(1) = -2.f + 4.f * frand(); Xw(2) = 8.f + 30.f * frand();
            for (int k = 0; k < NKF; ++k) {
                const float xl = Xw(0) - 0.3f * k, yl = Xw(1), zl = Xw(2) - 1.0f * k;
                lm->addObservationOnKeyframe(Pixel(fx * xl / zl + cx, fy * yl / zl + cy), lm_f[k]);
                lm_f[k]->addRelatedLandmark(lm);
            }
            Point Xn = Xw; Xn(2) *= 1.f + 0.02f * (frand() - 0.5f);
            lm->set3DPoint(Xn);
            lm_l.push_back(lm);
        }
        for (int k = 0; k < NKF; ++k) kwin->addNewKeyframe(lm_f[k]);
        const float err_before = std::fabs(lm_f[3]->getPose()(2, 3) - 3.0f) + std::fabs(lm_f[3]->getPose()(0, 3) - 0.9f);
        CHECK(me_mono.localBundleAdjustmentSparseSolver(kwin, camc), "mono local-BA driver runs");
        const float err_after = std::fabs(lm_f[3]->getPose()(2, 3) - 3.0f) + std::fabs(lm_f[3]->getPose()(0, 3) - 0.9f);
        CHECK(lm_l[0]->isBundled() && err_after <= err_before + 1e-3f, "mono driver bundles and does not diverge");
    }

    // ---- two FeatureTracker objects interleaved on the shared context: neither may see the other's upload as its own
    {
        std::vector<unsigned char> c0, c1;
        render(c0, W, H, 7.f, 3.f);
        render(c1, W, H, 7.f + sx, 3.f + sy);
        cv::Mat jmg0(H, W, c0.data()), jmg1(H, W, c1.data());
        FeatureTracker fa, fb;
        PixelVec pa, pb, pa2; MaskVec ma, mb, ma2;
        fa.track(img0, img1, pts0, 21, 3, 30.f, pa, ma);
        fb.track(jmg0, jmg1, pts0, 21, 3, 30.f, pb, mb);          // evicts fa's images from the shared slots
        fb.track(jmg1, jmg0, pts0, 21, 3, 30.f, pb, mb);
        fa.track(img0, img1, pts0, 21, 3, 30.f, pa2, ma2);         // must re-upload, not hit a stale fingerprint
        bool eq = true;
        for (size_t i = 0; i < n; ++i) eq &= pa[i].x == pa2[i].x && pa[i].y == pa2[i].y && ma[i] == ma2[i];
        CHECK(eq, "interleaved FeatureTracker objects do not alias each other's image slots");
        // du0 / dv0: the true Sobel images are accepted, anything else is refused
        std::vector<float> du((size_t)W * H), dv((size_t)W * H);
        auto pxl = [&](int xx, int yy) -> float {
            if (xx < 0) xx = -xx;
            if (xx >= W) xx = 2 * W - 2 - xx;
            if (yy < 0) yy = -yy;
            if (yy >= H) yy = 2 * H - 2 - yy;
            return (float)b0[(size_t)yy * W + xx];
        };
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            du[(size_t)y * W + x] = (pxl(x + 1, y - 1) - pxl(x - 1, y - 1)) + 2.f * (pxl(x + 1, y) - pxl(x - 1, y)) + (pxl(x + 1, y + 1) - pxl(x - 1, y + 1));
            dv[(size_t)y * W + x] = (pxl(x - 1, y + 1) - pxl(x - 1, y - 1)) + 2.f * (pxl(x, y + 1) - pxl(x, y - 1)) + (pxl(x + 1, y + 1) - pxl(x + 1, y - 1));
        }
        cv::Mat du0(H, W, du.data()), dv0(H, W, dv.data());
        PixelVec p6 = p1; MaskVec m6;
        fa.trackWithScale(img0, du0, dv0, img1, pts0, scale, p6, m6);
        bool eq5 = true;
        for (size_t i = 0; i < n; ++i) eq5 &= p6[i].x == p5[i].x && p6[i].y == p5[i].y && m6[i] == m5[i];
        CHECK(eq5, "trackWithScale with the true Sobel images == with empty derivative arguments");
        du[0] += 3.f;                              // pixel (0, 0) is part of the verified sample
        CHECK(throws_with([&] { PixelVec q = p1; MaskVec mm; fa.trackWithScale(img0, du0, dv0, img1, pts0, scale, q, mm); }, "cv::Sobel(img0"),
              "trackWithScale refuses derivative images that are not the Sobel of img0");
    }

    vo_b200::release_shared_context();
    printf("SHIM_OK features=%zu tracked=%zu\n", n, good);
    return 0;
}
