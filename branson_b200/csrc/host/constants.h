// constants.h -- physical constants and enumerations with the reference's names and values (src/constants.h:15-35).
#pragma once

namespace branson {
namespace Constants {
constexpr double pi = 3.1415926535897932384626433832795;
constexpr double c = 299.792458;  // cm / shake
constexpr double a = 0.01372;     // GJ / cm^3 / keV^4
constexpr double cutoff_fraction = 0.01;

enum bc_type { REFLECT, VACUUM, ELEMENT, SOURCE, PROCESSOR };
enum dir_type { X_NEG, X_POS, Y_NEG, Y_POS, Z_NEG, Z_POS };
enum event_type : unsigned char { EXIT, PASS, CENSUS, SCATTER, KILLED, BOUND };
enum { PARTICLE_PASS, REPLICATED };
enum { AOS, SOA };
enum { HISTORY, EVENT };
enum { NO_DECOMP, METIS, CUBE };
}  // namespace Constants
}  // namespace branson
