// Stub for the vendored tinyparser-mitsuba / tinyxml2 sources, which include Lumen precompiled header.
