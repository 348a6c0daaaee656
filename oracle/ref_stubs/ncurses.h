// No-op stand-in for the ncurses calls in src/galileo-sdr.cpp:432,619-652 (status table, -v only).
#pragma once
#define A_REVERSE 0
static inline void *initscr(void) { return 0; }
static inline int endwin(void) { return 0; }
static inline int clear(void) { return 0; }
static inline int refresh(void) { return 0; }
static inline int attron(int) { return 0; }
static inline int attroff(int) { return 0; }
static inline int printw(const char *, ...) { return 0; }
