// flye_b200 host mirror — the slice of the reference's Logger (src/common/logger.h:12-112) the sequence layer uses:
// Logger::get().info()/debug()/warning()/error() << ... ; messages go to stderr, debug only when enabled.
#pragma once
#include <iostream>
#include <sstream>
#include <string>

class Logger {
public:
    static Logger& get() {
        static Logger l;
        return l;
    }
    void setDebugging(bool on) { _debug = on; }
    class Line {
    public:
        Line(const char* tag, bool on) : _on(on) { if (_on) _ss << tag; }
        Line(Line&& o) : _ss(std::move(o._ss)), _on(o._on) { o._on = false; }
        ~Line() { if (_on) std::cerr << _ss.str() << std::endl; }
        template <class T> Line& operator<<(const T& v) { if (_on) _ss << v; return *this; }
    private:
        std::ostringstream _ss;
        bool _on;
    };
    Line info() { return Line("INFO: ", true); }
    Line warning() { return Line("WARNING: ", true); }
    Line error() { return Line("ERROR: ", true); }
    Line debug() { return Line("DEBUG: ", _debug); }
private:
    bool _debug = false;
};
