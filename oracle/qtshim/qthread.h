// TEST INFRASTRUCTURE: minimal stand-in for QThread as used by NetworkThread (B/network.h:71,531-555) under _NO_GUI.
#pragma once
class QThread {
  public:
    enum Priority { HighestPriority };
    virtual ~QThread() {}
    virtual void run() {}
    void start(Priority = HighestPriority) { run(); }
    void wait() {}
};
