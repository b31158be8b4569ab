// C++ facade tests: the reference's own test cases (test/testContactModel.cpp, test/UnitTest.cpp) written against
// include/eslam/*.hpp, i.e. through the reference's class names, running on the GPU via libeslam_gpu.so.
// FakeMLSAccess (test/testContactModel.cpp:8-38) becomes a 2x2 DenseGrid with 1000 m cells (same 4-quadrant rule).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <eslam/EmbodiedSlamFilter.hpp>

using namespace eslam;
using namespace odometry;

static int failures = 0;

// compile-time check of the snapshot type (src/PoseParticle.hpp:88-114): filled the way the reference's task fills it
static inline PoseDistribution snapshot(EmbodiedSlamFilter& f, const base::Quaterniond& q, const BodyContactState& bs) {
    PoseDistribution pd;
    pd.time = base::Time::fromMicroseconds(1);
    pd.particles.assign(f.getParticles().begin(), f.getParticles().end());
    pd.orientation = q; pd.bodyState = bs;
    return pd;
}

#define CHECK(cond) do { if (!(cond)) { std::printf("CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)
#define CHECK_SMALL(v, tol) CHECK(std::fabs(v) < (tol))
#define CHECK_CLOSE(a, b, pct) CHECK(std::fabs((a) - (b)) <= (pct) / 100.0 * std::fabs(b))

static DenseGrid quadrantGrid(const double* z, const double* stddev, const bool* res = 0) {
    DenseGrid g(2, 2, 1000.0, 1000.0, -1000.0, -1000.0);
    for (int iy = 0; iy < 2; iy++) for (int ix = 0; ix < 2; ix++) {
        const int i = 2 * iy + ix;
        g.set(ix, iy, (res && !res[i]) ? std::nanf("") : (float)z[i], (float)stddev[i]);
    }
    return g;
}

static BodyContactState fourPoints(const double (*p)[3], const float* contact, const int* group) {
    BodyContactState state; state.points.resize(4);
    for (int i = 0; i < 4; i++) {
        state.points[i].position = base::Vector3d(p[i][0], p[i][1], p[i][2]);
        state.points[i].contact = contact[i]; state.points[i].slip = 0; state.points[i].groupId = group[i];
    }
    return state;
}

static void test_updatePose_nogroup() {                    // test/testContactModel.cpp:128-190
    const double pts[4][3] = {{-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0}};
    const float contact[4] = {0.5f, 0.5f, 0.5f, 0.5f}; const int group[4] = {-1, -1, -1, -1};
    BodyContactState state = fourPoints(pts, contact, group);
    ContactModel model;
    ContactModelConfiguration cfg; cfg.contactPointRadius = 0.0;     // the reference's expectations predate the 0.01 radius
    model.setConfiguration(cfg);
    model.setContactPoints(state, base::Quaterniond::Identity());
    {
        double z[4] = {0, 0, 0, 0}, stddev[4] = {1, 1, 1, 1};
        model.setMap(quadrantGrid(z, stddev));
        CHECK(model.evaluatePose(base::Affine3d::Identity(), 1));
        CHECK(model.getContactPoints().size() == 4);
        for (int i = 0; i < 4; i++) {
            const ContactPoint& c = model.getContactPoints()[i];
            CHECK_SMALL((c.point - base::Vector3d(pts[i][0], pts[i][1], z[i])).norm(), 1e-6);
        }
        CHECK_SMALL(model.getZDelta(), 1e-6);
        CHECK_CLOSE(model.getZVar(), 0.5, 1e-6);
        CHECK_CLOSE(model.getWeight(), 1.0, 1e-6);
    }
    {
        double z[4] = {0, -0.12, -0.12, -0.12}, stddev[4] = {1, 1e9, 1e9, 1e9};
        model.setMap(quadrantGrid(z, stddev));
        CHECK(model.evaluatePose(base::Affine3d::Identity(), 1));
        CHECK_SMALL(model.getZDelta(), 1e-6);
        CHECK_CLOSE(model.getZVar(), 2.0, 1e-6);
        CHECK_CLOSE(model.getWeight(), 1.0, 1e-6);
    }
}

static void test_lowest_points() {                         // test/testContactModel.cpp:193-279
    const double pts[4][3] = {{-1, -1, 0.1}, {1, -1, -0.1}, {-1, 1, 0.1}, {1, 1, -0.1}};
    const float contact[4] = {1, 2, 3, 4};
    {
        const int group[4] = {-1, -1, -1, -1};
        ContactModel model; model.setContactPoints(fourPoints(pts, contact, group), base::Quaterniond::Identity());
        CHECK(model.getLowestPointPerGroup().size() == 4);
        model.updateContactStateUsingLowestPointHeuristic();
        for (int i = 0; i < 4; i++) CHECK(model.getContactState().points[i].contact == contact[i]);
    }
    {
        const int group[4] = {0, 0, 1, 1};
        ContactModel model; model.setContactPoints(fourPoints(pts, contact, group), base::Quaterniond::Identity());
        const std::vector<base::Vector3d>& low = model.getLowestPointPerGroup();
        CHECK(low.size() == 2);
        CHECK(low[0] == base::Vector3d(1, -1, -0.1) && low[1] == base::Vector3d(1, 1, -0.1));
        for (int i = 0; i < 4; i++) CHECK(model.getContactState().points[i].contact == contact[i]);
        model.updateContactStateUsingLowestPointHeuristic();
        const float expected[4] = {0, 1, 0, 1};
        for (int i = 0; i < 4; i++) CHECK(model.getContactState().points[i].contact == expected[i]);
    }
}

static void test_mapAbsence_group() {                      // test/testContactModel.cpp:326-362 (minContacts <= 1)
    const double pts[4][3] = {{-1, -1, 0.1}, {1, -1, -0.1}, {-1, 1, 0.1}, {1, 1, -0.1}};
    const float u = base::unknown<float>(); const float contact[4] = {u, u, u, u}; const int group[4] = {0, 0, 1, 1};
    ContactModel model;
    ContactModelConfiguration cfg; cfg.minContacts = 1; model.setConfiguration(cfg);
    model.setContactPoints(fourPoints(pts, contact, group), base::Quaterniond::Identity());
    double z[4] = {-0.1, -0.1, -0.1, -0.1}, stddev[4] = {1e9, 1, 1e9, 1}; bool res[4] = {true, true, true, false};
    model.setMap(quadrantGrid(z, stddev, res));
    CHECK(model.evaluatePose(base::Affine3d::Identity(), 1));
    CHECK(model.getContactPoints().size() == 1);
    bool threw = false;
    try { model.evaluatePose(base::Affine3d::Identity(), 0); } catch (std::runtime_error& e) { threw = std::string(e.what()).find("zero measurement variance") != std::string::npos; }
    CHECK(threw);                                           // src/ContactModel.cpp:122-123
}

static DenseGrid rollingTerrain(int n) {
    DenseGrid g(n, n, 0.05, 0.05, -n * 0.025, -n * 0.025);
    for (int iy = 0; iy < n; iy++) for (int ix = 0; ix < n; ix++) {
        const double x = g.offsetX + (ix + 0.5) * 0.05, y = g.offsetY + (iy + 0.5) * 0.05;
        g.set(ix, iy, (float)(0.25 * std::sin(2 * M_PI * x / 5.0) * std::cos(2 * M_PI * y / 7.0)), 0.02f);
    }
    return g;
}

static BodyContactState wheels() {
    BodyContactState s;
    const double c[4][2] = {{0.25, 0.25}, {0.25, -0.25}, {-0.25, 0.25}, {-0.25, -0.25}};
    for (int g = 0; g < 4; g++) for (int j = 0; j < 5; j++) {
        BodyContactPoint p; const double a = 2 * M_PI * j / 5 - M_PI / 2 + 0.3 * g;
        p.position = base::Vector3d(c[g][0] + 0.18 * std::cos(a), c[g][1], 0.18 * std::sin(a));
        p.contact = 0.5f; p.slip = 0; p.groupId = g; s.points.push_back(p);
    }
    return s;
}

static void test_filter_steps() {                          // PoseEstimator / EmbodiedSlamFilter driven like the Rock task does
    eslam::Configuration conf; conf.particleCount = 2000;
    {   // "No environment attached." (src/PoseEstimator.cpp:259-260)
        FootContact odo; PoseEstimator pe(odo, conf);
        pe.init(100, base::Pose2D(), base::Pose2D(base::Vector2d(0.1, 0.1), 0.1), 0.19, 1.001);
        bool threw = false;
        try { pe.update(wheels(), base::Quaterniond::Identity(), std::vector<terrain_estimator::TerrainClassification>()); }
        catch (std::runtime_error& e) { threw = std::string(e.what()) == "No environment attached."; }
        CHECK(threw);
    }
    {   // init without a map throws like the reference (src/EmbodiedSlamFilter.cpp:104)
        EmbodiedSlamFilter f(odometry::Configuration(), conf);
        bool threw = false;
        try { f.init(NULL, base::Pose()); } catch (std::runtime_error&) { threw = true; }
        CHECK(threw);
    }
    EmbodiedSlamFilter f(odometry::Configuration(), conf);
    DenseGrid map = rollingTerrain(400);
    f.init(&map, base::Pose(base::Vector3d(0, 0, 0.19), base::Quaterniond::Identity()));
    CHECK(f.getParticles().size() == 2000);
    for (size_t i = 0; i < 10; i++) CHECK(f.getParticles()[i].weight == 0.0);     // ctor leaves weight = 0 (quirk Q9)
    std::vector<terrain_estimator::TerrainClassification> ltc;
    BodyContactState bs = wheels();
    int updates = 0;
    for (int step = 0; step < 40; step++) {
        const double yaw = 0.005 * step;
        base::Affine3d body2odo = base::Affine3d::fromTranslationYaw(base::Vector3d(0.02 * step, 0, 0.19), yaw);
        if (f.update(body2odo, bs, ltc)) updates++;
        CHECK(f.getParticles().size() == 2000);               // test/UnitTest.cpp:117: resampling keeps the particle count
    }
    CHECK(updates >= 5 && updates <= 12);                      // measurementThreshold gates the update (src/EmbodiedSlamFilter.cpp:360)
    double sum = 0; for (size_t i = 0; i < 2000; i++) sum += f.getParticles()[i].weight;
    const size_t best = f.getBestParticleIndex();
    for (size_t i = 0; i < 2000; i++) CHECK(!(f.getParticles()[i].weight > f.getParticles()[best].weight));
    base::Affine3d c = f.getCentroid();
    CHECK(std::fabs(c.translation().x() - 0.78) < 0.3 && std::fabs(c.translation().y()) < 0.3);
    CHECK(std::fabs(f.getFilter().getWeightsSum() - 1.0) < 1e-9);   // getCentroid normalised the weights (src/PoseEstimator.cpp:356)
    // mutable getParticles(): edits reach the device before the next step (src/ParticleFilter.hpp:150-153)
    f.getParticles()[7].weight = 5.0;
    CHECK(f.getBestParticleIndex() == 7);
    (void)sum;
}

static void test_per_particle_maps() {                     // init(..., useSharedMap = false): src/EmbodiedSlamFilter.cpp:131-134,179-232
    eslam::Configuration conf; conf.particleCount = 1500; conf.minEffective = 1501; conf.useVisualUpdate = true;
    EmbodiedSlamFilter f(odometry::Configuration(), conf);
    DenseGrid map = rollingTerrain(400);
    f.init(&map, base::Pose(base::Vector3d(0, 0, 0.19), base::Quaterniond::Identity()), false);
    CHECK(!f.getFilter().usesSharedMap());
    CHECK(f.getParticles().size() == 1500);
    std::vector<terrain_estimator::TerrainClassification> ltc;
    BodyContactState bs = wheels();
    // a scan 0.2 m in front of the body, 8 x 6 cells of 0.05 m, heights relative to the body (z = 0.19 above the ground)
    DenseGrid scan(8, 6, 0.05, 0.05, 0.2, -0.15);
    for (int iy = 0; iy < 6; iy++) for (int ix = 0; ix < 8; ix++) scan.set(ix, iy, -0.19f, 0.02f);
    int mapped = 0;
    for (int step = 0; step < 12; step++) {
        base::Affine3d body2odo = base::Affine3d::fromTranslationYaw(base::Vector3d(0.03 * step, 0, 0.19), 0.0);
        f.update(body2odo, bs, ltc);                          // contact update: resamples every time (minEffective > N): maps travel
        if (f.update(body2odo, scan)) mapped++;               // mapping update past mappingThreshold: match weight + merge
        CHECK(f.getParticles().size() == 1500);
    }
    CHECK(mapped >= 3 && mapped <= 6);                         // mappingThreshold gates the mapping update (:325; 0.087 m with the swapped arguments)
    std::vector<float> cells((size_t)1500 * 16 * 16 * 2); std::vector<int32_t> origin(2 * 1500);
    CHECK(eslam_gpu_get_patches(f.getFilter().context(), cells.data(), origin.data()) == ESLAM_OK);
    size_t written = 0; for (size_t i = 0; i < cells.size(); i += 2) written += cells[i] == cells[i];
    CHECK(written > 1500 * 20);                                // every particle's own map holds scan and contact heights
    // the logging snapshot: particles + one-component mixture around the cloud (src/PoseParticle.hpp:88-114)
    PoseDistribution pd = f.getPoseDistribution(base::Time::fromMicroseconds(42));
    CHECK(pd.particles.size() == 1500 && pd.gmm.params.size() == 1 && pd.bodyState.points.size() == bs.points.size());
    CHECK(std::fabs(pd.gmm.params[0].mean.x() - 0.33) < 0.3 && pd.gmm.params[0].cov.m[0][0] > 0 && pd.gmm.params[0].cov.m[1][1] > 0);
    // resample_stratified(samples) / resample_multinomial(samples) (src/ParticleFilter.hpp:85-147) change the set size
    f.getFilter().resample_stratified(1000);
    CHECK(f.getParticles().size() == 1000);
    f.getFilter().resample_multinomial(800);
    CHECK(f.getParticles().size() == 800);
    CHECK(std::fabs(f.getParticles()[3].weight - 1.0 / 1000) < 1e-15);   // :140
}

static void test_surface_hash() {                          // init(env, pose, true, hashConfig with useHash): src/EmbodiedSlamFilter.cpp:93-98,109-114
    eslam::Configuration conf; conf.particleCount = 3000; conf.minEffective = 3001;
    SurfaceHashConfig hc; hc.useHash = true; hc.period = 3;
    EmbodiedSlamFilter f(odometry::Configuration(), conf);
    DenseGrid map = rollingTerrain(200);
    f.init(&map, base::Pose(base::Vector3d(0, 0, 0.19), base::Quaterniond::Identity()), true, hc);
    CHECK(f.getParticles().size() == 3000);
    double minx = 1e9, maxx = -1e9;
    for (size_t i = 0; i < 3000; i++) { minx = std::fmin(minx, f.getParticles()[i].position.x()); maxx = std::fmax(maxx, f.getParticles()[i].position.x()); }
    CHECK(maxx - minx > 5.0);                                 // seeded all over the 10 m map, not around the start pose
    CHECK(f.getParticles()[0].zSigma == 0.0 && f.getParticles()[0].weight == 0.0);
    std::vector<terrain_estimator::TerrainClassification> ltc;
    BodyContactState bs = wheels();
    for (int step = 0; step < 9; step++) {
        base::Affine3d body2odo = base::Affine3d::fromTranslationYaw(base::Vector3d(0.03 * step, 0, 0.19), 0.0);
        f.update(body2odo, bs, ltc);                          // project() re-seeds from the hash every 3rd call (:239-241)
        CHECK(f.getParticles().size() == 3000);
    }
    SurfaceParam sp; std::vector<base::Vector3d> pts;        // test/UnitTest.cpp:121-143 exercises fromPoints
    pts.push_back(base::Vector3d(0.25, 0, 0.05)); pts.push_back(base::Vector3d(-0.25, 0, -0.05));
    pts.push_back(base::Vector3d(0.25, -0.5, 0.05)); pts.push_back(base::Vector3d(-0.25, -0.5, -0.05));
    sp.fromPoints(pts);
    CHECK_CLOSE(sp.slope_x, 0.2, 1e-9); CHECK_SMALL(sp.slope_y, 1e-12);
}

static void test_z_position_estimate() {                   // ContactModel::updateZPositionEstimate as test/testMap.cpp:279-289 drives it
    const double pts[4][3] = {{-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0}};
    const float contact[4] = {0.5f, 0.5f, 0.5f, 0.5f}; const int group[4] = {-1, -1, -1, -1};
    ContactModel model;
    ContactModelConfiguration cfg; cfg.contactPointRadius = 0.0; model.setConfiguration(cfg);
    model.setContactPoints(fourPoints(pts, contact, group), base::Quaterniond::Identity());
    double z[4] = {0.05, 0.05, 0.05, 0.05}, stddev[4] = {0.1, 0.1, 0.1, 0.1};
    model.setMap(quadrantGrid(z, stddev));
    CHECK(model.evaluatePose(base::Affine3d::Identity(), 0.5));
    double zpos = 0.0, zvar = 0.5;
    CHECK(model.updateZPositionEstimate(zpos, zvar));
    CHECK(zpos > 0.0 && zpos < 0.05 + 1e-9);                   // pulled towards the ground the feet stand on
    CHECK(zvar < 0.5 && zvar > 0.0);
}

int main() {
    test_updatePose_nogroup();
    test_per_particle_maps();
    test_surface_hash();
    test_z_position_estimate();
    test_lowest_points();
    test_mapAbsence_group();
    test_filter_steps();
    if (failures) { std::printf("%d CHECK(s) failed\n", failures); return 1; }
    std::printf("facade tests ok\n");
    return 0;
}
