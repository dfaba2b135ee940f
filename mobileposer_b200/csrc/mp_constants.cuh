// The numbers the kernels bake in, in ONE place.  They mirror mobileposer_b200/config.py (which mirrors the reference's
// mobileposer/config.py:129-142, models/net.py:47-59 and the SMPL zero pose of articulate/model.py:77-92); mp_constants()
// (api.cu) hands them to the host and tests/test_constants.py checks both copies agree bit for bit.
#pragma once

// SMPL kinematic tree (smpl/basicmodel_m.pkl kintree_table), -1 = root
#define MP_SMPL_PARENT_INIT {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21}
// joint_set.reduced / joint_set.ignored (config.py:134-135)
#define MP_REDUCED_INIT {0, 1, 2, 3, 4, 5, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19}
#define MP_IGNORED_INIT {0, 7, 8, 10, 11, 20, 21, 22, 23}
// joint -> slot in `reduced` or -1 (derived; test_constants.py re-derives it)
#define MP_REDUCED_SLOT_INIT {0, 1, 2, 3, 4, 5, 6, -1, -1, 7, -1, -1, 8, 9, 10, 11, 12, 13, 14, 15, -1, -1, -1, -1}
// zero-pose joints J - J[0] as float32
#define MP_SMPL_J_ZERO_INIT                                                                                              \
    {{0.0f, 0.0f, 0.0f},                                                                                                \
     {0.058581352f, -0.08228004f, -0.017664082f},                                                                       \
     {-0.060309727f, -0.09051329f, -0.013542531f},                                                                      \
     {0.004439451f, 0.12440355f, -0.03838522f},                                                                         \
     {0.10203278f, -0.46874952f, -0.009627081f},                                                                        \
     {-0.103566356f, -0.47420114f, -0.018385574f},                                                                      \
     {0.008927891f, 0.26235995f, -0.0115648955f},                                                                       \
     {0.08724245f, -0.8956239f, -0.047055073f},                                                                         \
     {-0.08451081f, -0.8942467f, -0.052947246f},                                                                        \
     {0.0066633024f, 0.31839234f, -0.008709848f},                                                                       \
     {0.1282968f, -0.95590985f, 0.074987344f},                                                                          \
     {-0.11935069f, -0.95635235f, 0.07737604f},                                                                         \
     {-0.006726882f, 0.53002787f, -0.042177428f},                                                                       \
     {0.07836577f, 0.43239203f, -0.02760802f},                                                                          \
     {-0.076290354f, 0.4308647f, -0.032417234f},                                                                        \
     {0.0033863292f, 0.61896527f, 0.008232435f},                                                                        \
     {0.20128717f, 0.47759712f, -0.04665402f},                                                                          \
     {-0.18951866f, 0.47771794f, -0.040889304f},                                                                        \
     {0.45661905f, 0.4619481f, -0.06960051f},                                                                           \
     {-0.44964615f, 0.46334866f, -0.07215803f},                                                                         \
     {0.7223283f, 0.4746462f, -0.07697524f},                                                                            \
     {-0.7187546f, 0.47014236f, -0.0781848f},                                                                           \
     {0.80901885f, 0.46401018f, -0.09256954f},                                                                          \
     {-0.80750835f, 0.4614908f, -0.088291876f}}
// zero-pose feet J[10], J[11] (net.py:47-48,59): rows 10 and 11 of the table above
#define MP_FEET_INIT {0.1282968f, -0.95590985f, 0.074987344f, -0.11935069f, -0.95635235f, 0.07737604f}

namespace mp {
constexpr double kFloorY = -0.9563523530960083;   // float32 min(J[10].y, J[11].y) widened (net.py:49)
constexpr float kGravityVel = -0.018f;            // joint_set.gravity_velocity (config.py:131)
constexpr float kVelDiv = 15.0f;                  // datasets.fps / amass.vel_scale (net.py:141)
constexpr float kProbLo = 0.5f, kProbHi = 0.9f;   // prob_threshold (net.py:53)
}  // namespace mp
