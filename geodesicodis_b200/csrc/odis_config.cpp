// input.in reader and derived constants (see odis_config.h). Citations are to
// /root/reference/src/globals.cpp unless another file is named.
#include "odis_config.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "odis_sphere.h"

namespace odis {

void Config::reg(const char* key, ConfigEntry::Type t) {
    ConfigEntry e;
    e.type = t;
    entries_[key] = e;
    order_.push_back(key);
}

Config::Config() {
    // registration order and key strings: globals.cpp:58-203
    const char* doubles[] = {"radius", "angular velocity", "k2", "h2", "love reduction factor", "ocean thickness",
                             "shell thickness", "surface gravity", "semimajor axis", "eccentricity", "obliquity",
                             "time step", "converge", "friction coefficient"};
    for (const char* k : doubles) reg(k, ConfigEntry::DOUBLE);
    reg("sh degree", ConfigEntry::INT);
    reg("latitude spacing", ConfigEntry::DOUBLE);
    reg("longitude spacing", ConfigEntry::DOUBLE);
    reg("geodesic grid level", ConfigEntry::INT);
    reg("orbital period", ConfigEntry::DOUBLE);
    reg("advection", ConfigEntry::BOOL);
    reg("simulation end time", ConfigEntry::DOUBLE);
    reg("fourier a20", ConfigEntry::DOUBLE);
    reg("fourier a22", ConfigEntry::DOUBLE);
    reg("fourier b22", ConfigEntry::DOUBLE);
    reg("fourier freq", ConfigEntry::INT);
    const char* strings[] = {"potential", "friction type", "surface type", "solver type", "initial conditions"};
    for (const char* k : strings) reg(k, ConfigEntry::STRING);
    const char* bools[] = {"dissipation avg output", "kinetic avg output", "work flux output", "displacement output",
                           "velocity cartesian output", "velocity output", "pressure output", "dissipation output",
                           "kinetic output", "dummy1 output", "dummy2 output", "sh coefficient output"};
    for (const char* k : bools) reg(k, ConfigEntry::BOOL);
    reg("output time", ConfigEntry::INT);
    reg("total iterations", ConfigEntry::INT);
    reg("core number", ConfigEntry::INT);
    reg("grav coeff file", ConfigEntry::STRING);
    reg("forcing coeff file", ConfigEntry::STRING);
    reg("fourier coeff file", ConfigEntry::STRING);
    reg("rbf epsilon", ConfigEntry::DOUBLE);

    // Titan defaults: globals.cpp:479-571
    entries_["core number"].i = 1;
    entries_["advection"].b = false;
    entries_["total iterations"].i = 1000;
    entries_["radius"].d = 2574.73e3;
    entries_["k2"].d = .120;
    entries_["h2"].d = .2;
    entries_["love reduction factor"].d = 0.920012;
    entries_["ocean thickness"].d = 400;
    entries_["surface gravity"].d = 1.35;
    entries_["semimajor axis"].d = 1221.87e6;
    entries_["eccentricity"].d = 0.0288;
    entries_["obliquity"].d = 0.32 * kPi / 180.;
    entries_["friction coefficient"].d = 2.28e-7;
    entries_["sh degree"].i = 10;
    entries_["latitude spacing"].d = 45;
    entries_["longitude spacing"].d = 90;
    entries_["geodesic grid level"].i = 1;
    entries_["time step"].d = 40.;
    entries_["converge"].d = 1e-7;
    entries_["angular velocity"].d = 4.56e-6;
    entries_["orbital period"].d = 2 * kPi / 4.56e-6;
    entries_["simulation end time"].d = 80.;
    entries_["potential"].s = "ECC";
    entries_["friction type"].s = "QUADRATIC";
    entries_["initial conditions"].s = "NONE";
    entries_["dissipation avg output"].b = true;
    entries_["kinetic avg output"].b = false;
    entries_["work flux output"].b = false;
    entries_["output time"].i = 1;
    entries_["rbf epsilon"].d = 0.25;
}

const ConfigEntry* Config::find(const std::string& key) const {
    auto it = entries_.find(key);
    return it == entries_.end() ? nullptr : &it->second;
}
double Config::get_double(const std::string& key) const { return entries_.at(key).d; }
int Config::get_int(const std::string& key) const { return entries_.at(key).i; }
bool Config::get_bool(const std::string& key) const { return entries_.at(key).b; }
const std::string& Config::get_string(const std::string& key) const { return entries_.at(key).s; }

std::vector<std::string> Config::unassigned() const {
    std::vector<std::string> r;
    for (const auto& k : order_)
        if (!entries_.at(k).assigned) r.push_back(k);
    return r;
}

int Config::set(const std::string& key_in, const std::string& value_text) {
    std::string key = key_in;
    std::transform(key.begin(), key.end(), key.begin(), ::tolower);          // :360
    auto it = entries_.find(key);
    if (it == entries_.end()) return -1;                                      // unknown keys are ignored (:369-413)
    ConfigEntry& e = it->second;
    std::stringstream value(value_text);                                      // :364
    switch (e.type) {
        case ConfigEntry::DOUBLE: { double v = e.d; value >> v; e.d = v; break; }            // :377-380
        case ConfigEntry::INT: { int v = e.i; value >> v; e.i = v; break; }                  // :386-387
        case ConfigEntry::BOOL: {                                                            // :393-396
            std::string s; value >> s;
            if (s == "false") last_bool_ = false;
            else if (s == "true") last_bool_ = true;
            e.b = last_bool_;
            break;
        }
        case ConfigEntry::STRING: { std::string s; value >> s; e.s = s; break; }             // :401-402
    }
    e.assigned = true;
    return 0;
}

int Config::load(const std::string& dir, std::string& err) {
    run_dir = dir;
    std::ifstream in(dir + "/input.in", std::ifstream::in);
    if (!in.is_open()) {
        err = "Unable to open 'input.in' file.";                              // :471
        return -1;
    }
    std::string key, val, comment;
    while (std::getline(in >> std::ws, key, ';')) {                           // :355
        std::getline(in >> std::ws, val, ';');                                // :361
        set(key, val);
        std::getline(in, comment, ';');                                       // :417
    }
    loaded_from_file_ = true;
    return 0;
}

int Config::finalize(std::string& err) {
    ConfigEntry& theta = entries_["obliquity"];
    ConfigEntry& period = entries_["orbital period"];
    ConfigEntry& omega = entries_["angular velocity"];
    if (loaded_from_file_) {
        theta.d = theta.d * kPi / 180.;                                       // :424 (degrees in the file)
        period.d = 2. * kPi / omega.d;                                        // :427
        out_tags.clear();                                                     // :431-440
        const char* tag_keys[] = {"velocity output", "velocity cartesian output", "displacement output", "pressure output",
                                  "dissipation output", "kinetic output", "dissipation avg output", "kinetic avg output",
                                  "dummy1 output", "dummy2 output"};
        for (const char* k : tag_keys)
            if (entries_[k].b) out_tags.push_back(k);
    }
    // the period is forced to an even whole number of seconds and the spin rate follows it: :209-214
    period.d = 2. * kPi / omega.d;
    const int int_time = (int)std::round(period.d / 2) * 2;
    period.d = (double)int_time;
    omega.d = 2 * kPi / (double)int_time;

    const std::string& fr = entries_["friction type"].s;                      // :233-239
    if (fr == "LINEAR") fric_type = LINEAR;
    else if (fr == "QUADRATIC") fric_type = QUADRATIC;
    else { err = "ERROR: NO DRAG MODEL FOUND!"; return -2; }

    const std::string& so = entries_["solver type"].s;                        // :241-248
    if (so == "EULER") solver_type = EULER;
    else if (so == "AB3") solver_type = AB3;
    else if (so == "RK4") solver_type = RK4;
    else { err = "ERROR: NO SOLVER FOUND!"; return -3; }

    static const char* pot_names[] = {"OBLIQ", "OBLIQ_WEST", "OBLIQ_EAST", "ECC_RAD", "ECC_LIB", "ECC", "ECC_WEST", "ECC_EAST",
                                      "FULL", "FULL2", "TOTAL", "ECC_W3", "OBLIQ_W3", "PLANET", "PLANET_OBL", "GENERAL", "NONE"};
    const std::string& po = entries_["potential"].s;                          // :250-271
    int found = -1;
    for (int k = 0; k < 17; k++)
        if (po == pot_names[k]) found = k;
    if (found < 0) { err = "ERROR: NO POTENTIAL FORCING FOUND!"; return -4; }
    tide_type = (Potential)found;

    static const char* surf_names[] = {"FREE", "FREE_LOADING", "LID_LOVE", "LID_MEMBR", "LID_NUM", "LID_INF"};
    const std::string& su = entries_["surface type"].s;                       // :273-283
    found = -1;
    for (int k = 0; k < 6; k++)
        if (su == surf_names[k]) found = k;
    if (found < 0) { err = "ERROR: NO OCEAN SURFACE BOUNDARY CONDITION FOUND!"; return -5; }
    surface_type = (Surface)found;

    const std::string& in = entries_["initial conditions"].s;                 // :285-292
    if (in == "NONE") initial_condition = INIT_NONE;
    else if (in == "LOAD") initial_condition = INIT_LOAD;
    else if (in == "ANALYTICAL") initial_condition = INIT_ANALYTICAL;
    else { err = "ERROR: INITIAL CONDITION MUST BE 0 (none), 1 (load from file), or 2 (analytical)!"; return -6; }

    // ---- surface boundary-condition factors: boundaryConditions.cpp:7-399 ----
    const int l_max = entries_["sh degree"].i;
    switch (surface_type) {
        case FREE:                                                            // boundaryConditions.cpp:18-25
        case LID_NUM:                                                         // :387-393
            entries_["love reduction factor"].d = 1.0 + entries_["k2"].d - entries_["h2"].d;
            break;
        case FREE_LOADING: {                                                  // :29-77 (Enceladus constants hard-wired there)
            const double ocean_den = 1000.0, bulk_den = 1609.22, rig = 40e9;
            const double g = entries_["surface gravity"].d, r = entries_["radius"].d;
            loading_factor.assign((size_t)l_max + 1, 0.0);
            for (int l = 0; l < l_max + 1; l++) {
                double eff_rig = (double)(2 * l * l + 4 * l + 3);
                eff_rig /= (double)l;
                eff_rig *= rig / (bulk_den * g * r);
                const double k_l = -(1.0 / (1.0 + eff_rig));
                const double h_l = -(1.0 / (1.0 + eff_rig)) * (2.0 * (double)l + 1.0) / 3.0;
                loading_factor[l] = (1.0 + k_l - h_l);
                loading_factor[l] *= 3. * ocean_den / ((2. * (double)l + 1.0) * bulk_den);
            }
            double eff_rig = (double)(2 * 2 * 2 + 4 * 2 + 3);
            eff_rig /= 2.0;
            eff_rig *= rig / (bulk_den * g * r);
            entries_["love reduction factor"].d = 1.0 + 1.5 / (1. + eff_rig) - 2.5 / (1. + eff_rig);
            break;
        }
        case LID_LOVE: {                                                      // :113-161,369-374
            entries_["radius"].d = entries_["radius"].d - entries_["shell thickness"].d;
            shell_factor_beta.assign((size_t)l_max + 1, 0.0);
            std::ifstream beta(run_dir + "/input_files/beta.txt", std::ifstream::in);
            if (beta.is_open()) {
                std::string line, val;
                int row = 1;
                while (std::getline(beta, line) && row <= l_max) {
                    std::istringstream ls(line);
                    std::getline(ls, val, '\t');
                    shell_factor_beta[row] = std::atof(val.c_str());
                    row++;
                }
            }
            for (int l = 0; l < l_max + 1; l++) shell_factor_beta[l] = 1.0 - shell_factor_beta[l];
            break;
        }
        case LID_MEMBR: {                                                     // :80-110 + membraneNuBeta, membraneConstants.cpp:9-143
            // Beuthe (2016) membrane shell, Enceladus constants hard-wired in the reference. Besides beta_l and the tidal
            // prefactor nu_2 it REPLACES g by the gravity at the ocean top and the radius by the ocean-top radius.
            const double G = 6.67384e-11, pois_ratio = 0.5, rigid_shell = 3.5e9, rigid_core = 40e9;
            const double shell_thickness = entries_["shell thickness"].d, ocean_thickness = entries_["ocean thickness"].d;
            const double radius = entries_["radius"].d;
            const double radius_core = radius - (shell_thickness + ocean_thickness), radius_ocean = radius - (shell_thickness);
            const double mass_total = 1.08e20, den_ocean = 1000.0, den_shell = 940.0;
            const double grav_surf = G * mass_total / std::pow(radius, 2.0);
            const double vol_total = 4. / 3. * kPi * std::pow(radius, 3.0);
            const double vol_core = 4. / 3. * kPi * std::pow(radius_core, 3.0);
            const double vol_ocean = 4. / 3. * kPi * std::pow(radius_ocean, 3.0) - vol_core;
            const double vol_shell = vol_total - 4. / 3. * kPi * std::pow(radius_ocean, 3.0);
            const double den_bulk = mass_total / vol_total;
            const double mass_ocean = vol_ocean * den_ocean, mass_shell = vol_shell * den_shell;
            const double mass_core = mass_total - (mass_ocean + mass_shell);
            const double den_core = mass_core / vol_core;
            const double grav_core = G * mass_core / std::pow(radius_core, 2.0);
            const double grav_ocean = G * (mass_core + mass_ocean) / std::pow(radius_ocean, 2.0);
            shell_factor_beta.assign((size_t)l_max + 1, 0.0);
            std::vector<double> nu((size_t)l_max + 1, 0.0);
            for (int l = 0; l < l_max + 1; l++) {
                const double rigid_eff = (double)(2 * l * l + 4 * l + 3) / ((double)l) * rigid_core / (den_core * grav_core * radius_core);
                const double rigid_factor = 1. / (1. + rigid_eff);
                double kt = rigid_factor * 3. / ((double)(2 * (l - 1)));
                double ht = rigid_factor * (double)(2 * l + 1) / ((double)(2 * (l - 1)));
                if (l == 1) { kt = 0.0; ht = 0.0; }
                const double kl = -rigid_factor;
                const double hl = -rigid_factor * (double)(2 * l + 1) / 3.0;
                const double gam_tide = 1.0 + kt - ht;
                const double gam_load = 1.0 + kl - hl;
                const double x = (double)((l - 1) * (l + 2));
                const double bendRigidity = rigid_shell * std::pow(shell_thickness, 3.0) / (6. * (1. - pois_ratio));
                const double sprMembrane = 2. * x * (1. + pois_ratio) / (x + 1. + pois_ratio) * rigid_shell / (den_ocean * grav_surf * radius) *
                                           shell_thickness / radius;
                const double sprBending = std::pow(x, 2.0) * (x + 2.) / (x + 1. + pois_ratio) * bendRigidity /
                                          (den_ocean * grav_surf * std::pow(radius, 4.0));
                const double sprConst = sprMembrane + sprBending;
                const double xi = 3.0 / (2. * (double)l + 1.0) * (den_ocean / den_bulk);
                double dsprConst = 1. - std::pow((1. + xi * hl), 2.0) / (1. + xi * (ht - hl) * sprConst);
                dsprConst *= -sprConst;
                double dgam_tide = (1. + xi * hl) * ht / (1. + xi * (ht - hl) * sprConst);
                dgam_tide *= -sprConst;
                shell_factor_beta[(size_t)l] = 1. - xi * gam_load + sprConst + dsprConst;     // beta(l): used as it is (:99), not 1 - beta
                nu[(size_t)l] = gam_tide + dgam_tide;
            }
            entries_["surface gravity"].d = grav_ocean;                       // membraneConstants.cpp:135-136
            entries_["radius"].d = radius_ocean;
            if (l_max >= 2) entries_["love reduction factor"].d = nu[2];      // boundaryConditions.cpp:101 reads (*nu)(2)
            break;
        }
        case LID_INF:
            break;
    }
    return 0;
}

}  // namespace odis
