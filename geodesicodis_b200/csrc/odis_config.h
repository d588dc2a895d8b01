// Run configuration: the input.in reader and the derived constants the solver uses.
//
// Replaces the reference `Globals` object (/root/reference/include/globals.h:82-205,
// /root/reference/src/globals.cpp:35-323 ctor, :325-477 ReadGlobals, :479-571 SetDefault) and
// the surface-type post-processing in /root/reference/src/boundaryConditions.cpp:7-399.
// Same key strings, same `key; value; comment;` syntax, same defaults (Titan), same period
// rounding; grid sizes are NOT part of the configuration (they come from the grid file).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace odis {

// enum orders follow include/globals.h:45-80 so integer values agree with the reference
enum Friction { LINEAR, QUADRATIC };
enum Surface { FREE, FREE_LOADING, LID_LOVE, LID_MEMBR, LID_NUM, LID_INF };
enum SolverKind { EULER, AB3, RK4 };
enum Potential {
    OBLIQ, OBLIQ_WEST, OBLIQ_EAST, ECC_RAD, ECC_LIB, ECC, ECC_WEST, ECC_EAST, FULL, FULL2, TOTAL,
    ECC_W3, OBLIQ_W3, PLANET, PLANET_OBL, GENERAL, NONE
};
enum InitKind { INIT_NONE, INIT_LOAD, INIT_ANALYTICAL };

struct ConfigEntry {
    enum Type { DOUBLE, INT, BOOL, STRING } type;
    double d = 0.0;
    int i = 0;
    bool b = false;
    std::string s;
    bool assigned = false;      // seen in input.in (globals.cpp:382 "Added")
};

class Config {
public:
    Config();                                        // registers keys + Titan defaults

    // Parse <run_dir>/input.in (globals.cpp:325-477). Returns 0 or <0 (error text in err).
    int load(const std::string& run_dir, std::string& err);
    // Set one key from text exactly as the file reader would. Returns 0, or -1 if the key is unknown.
    int set(const std::string& key, const std::string& value_text);
    // Derive everything the constructor derives after reading (globals.cpp:209-308) and
    // apply the surface boundary-condition factors. Must be called once before use.
    int finalize(std::string& err);

    double get_double(const std::string& key) const;
    int get_int(const std::string& key) const;
    bool get_bool(const std::string& key) const;
    const std::string& get_string(const std::string& key) const;
    bool has(const std::string& key) const { return entries_.count(key) != 0; }
    const ConfigEntry* find(const std::string& key) const;
    void set_double(const std::string& key, double v) { entries_[key].d = v; }
    void set_int(const std::string& key, int v) { entries_[key].i = v; }

    // keys in registration order (globals.cpp:58-203) and those never assigned
    const std::vector<std::string>& keys() const { return order_; }
    std::vector<std::string> unassigned() const;

    // ---- derived (valid after finalize) ----
    std::string run_dir;
    Friction fric_type = QUADRATIC;
    Surface surface_type = FREE;
    SolverKind solver_type = AB3;
    Potential tide_type = ECC;
    InitKind initial_condition = INIT_NONE;
    std::vector<std::string> out_tags;               // globals.cpp:431-440
    std::vector<double> shell_factor_beta;           // LID_LOVE: 1 - beta_l (boundaryConditions.cpp:123-161,369-374)
    std::vector<double> loading_factor;              // FREE_LOADING gamma_l (boundaryConditions.cpp:51-62)

private:
    void reg(const char* key, ConfigEntry::Type t);
    std::map<std::string, ConfigEntry> entries_;
    std::vector<std::string> order_;
    bool last_bool_ = true;                          // globals.cpp:341: valBool persists across keys
    bool loaded_from_file_ = false;
};

}  // namespace odis
