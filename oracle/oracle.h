/* oracle/oracle.h — TEST INFRASTRUCTURE ONLY: C entry points of liboracle.so (see api.cpp header). */
#ifndef ECHO_ORACLE_H
#define ECHO_ORACLE_H

#include "../include/echo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OracleScene OracleScene;

/* BxDF kinds for oracle_bxdf_batch (and the device mirror echo_b200_debug_bxdf_batch). params[11]:
 * [0..1] Trowbridge-Reitz alpha x/y (OrenNayar: [0] = roughness); RealFresnel: [2] etaAbove, [3] etaBelow;
 * ComplexFresnel: [2..4] etaAbove RGB, [5..7] etaBelow RGB, [8..10] extinction RGB. */
#define ORACLE_BXDF_LAMBERTIAN_REFLECTION 0
#define ORACLE_BXDF_LAMBERTIAN 1
#define ORACLE_BXDF_OREN_NAYAR 2
#define ORACLE_BXDF_SPECULAR_REFLECTION_REAL 3
#define ORACLE_BXDF_SPECULAR_REFLECTION_COMPLEX 4
#define ORACLE_BXDF_SPECULAR_TRANSMISSION 5
#define ORACLE_BXDF_SPECULAR_FRESNEL 6
#define ORACLE_BXDF_GLOSSY_REFLECTION_REAL 7
#define ORACLE_BXDF_GLOSSY_REFLECTION_COMPLEX 8
#define ORACLE_BXDF_GLOSSY_TRANSMISSION 9
#define ORACLE_BXDF_COATED_LAMBERTIAN 10 /* RealFresnel in [2..3], [4] = reflectance, [5..7] = albedo RGB */

#ifdef __cplusplus
}
#endif
#endif
